/* prosail_b200.h -- C ABI of libprosail_b200.so: batched PROSPECT-5/D -> LSM -> 4SAIL -> band means on B200 (sm_100a).
 *
 * The reference (ibaris/pyrism 0.0.5) is pure Python and has no FFI layer; its boundary for this path is the
 * constructor surface  pyrism.PROSPECT(...) / pyrism.LSM(...) / pyrism.SAIL(...) / pyrism.VolScatt(...).coef(...)
 * in pyrism/models/models.py.  Each entry point below names the reference interface it replaces.  The Python
 * drop-in classes in pyrism_b200/models.py bind these symbols with ctypes.
 *
 * Conventions
 *   - plain C: pointers, sizes and small POD structs; no C++/torch types cross the boundary.
 *   - return 0 on success, > 0 a cudaError_t, < 0 a PROSAIL_E_* code; prosail_last_error() gives the text
 *     (thread-local).  Nothing throws across the ABI.
 *   - on_host = 0: every data pointer is a DEVICE pointer on the handle's device, the call is asynchronous and
 *     ordered on `stream` (a cudaStream_t, NULL = the handle's own stream); the caller owns all buffers.
 *     on_host = 1: every data pointer is a HOST pointer; the library slices the work, overlaps H2D copies,
 *     kernels and D2H copies on two internal streams and returns when the results are in host memory.
 *   - angles are radians, already normalised as the reference's Kernel does (core/_core.py:136-146: negative
 *     zenith -> abs value and raa += pi).
 *   - spectra have PROSAIL_NBANDS = 2101 values (400..2500 nm); `pitch` is the row stride in elements
 *     (>= 2101; 2112 gives 128-byte aligned rows).
 *   - thread safety: the library has no global mutable state (the error text is thread-local), so any number of host
 *     threads may call it at the same time on DIFFERENT handles, several of which may live on the same device
 *     (each owns its constants, window list, workspace and streams: that is how the Python drop-in gives
 *     PROSPECT.select() / LSM.select() a window list of their own without touching the shared engine's).  Calls on
 *     ONE handle must not overlap in time: a handle carries the window list, the staged tables and one workspace.
 */
#ifndef PROSAIL_B200_H
#define PROSAIL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PROSAIL_NBANDS 2101
#define PROSAIL_NLIDF 18
#define PROSAIL_MAX_WINDOWS 64
#define PROSAIL_NGEOM 12 /* per-item scalars of prosail_out_t.geom / prosail_volscatt_f64 */

/* error codes (negative) */
#define PROSAIL_E_ARG (-1)      /* bad argument */
#define PROSAIL_E_STATE (-2)    /* tables / windows not set */
#define PROSAIL_E_NOGPU (-3)    /* no usable CUDA device: there is NO CPU fallback */

/* PROSPECT versions (reference: version='5' | 'D', models.py:915) */
#define PROSAIL_VERSION_5 0
#define PROSAIL_VERSION_D 1
/* LIDF types (reference: lidf_type='campbell' | 'verhoef', models.py:560-565) */
#define PROSAIL_LIDF_CAMPBELL 0
#define PROSAIL_LIDF_VERHOEF 1

/* output planes: index into prosail_out_t.plane[] and bit index of mean_mask */
enum {
    PROSAIL_O_RSOT = 0, /* BRF  = SAIL.BRF.ref  (models.py:575, rsot :726) */
    PROSAIL_O_RDDT,     /* BHR  = SAIL.BHR.ref  (rddt :721) */
    PROSAIL_O_RSDT,     /* DHR  = SAIL.DHR.ref  (rsdt :722) */
    PROSAIL_O_RDOT,     /* HDR  = SAIL.HDR.ref  (rdot :723) */
    PROSAIL_O_RSO,      /* SAIL.canopy.BRF (rso :710) */
    PROSAIL_O_RDD,      /* SAIL.canopy.BHR */
    PROSAIL_O_TDD,      /* SAIL.canopy.BHT */
    PROSAIL_O_RSD,      /* SAIL.canopy.DHR */
    PROSAIL_O_TSD,      /* SAIL.canopy.DHT */
    PROSAIL_O_RDO,      /* SAIL.canopy.HDR */
    PROSAIL_O_TDO,      /* SAIL.canopy.HDT */
    PROSAIL_O_KE,       /* SAIL.ke = m (models.py:601-602) */
    PROSAIL_O_LEAF_R,   /* PROSPECT.ks (leaf reflectance, models.py:1065) */
    PROSAIL_O_LEAF_T,   /* PROSPECT.kt (leaf transmittance, models.py:1064) */
    PROSAIL_O_RHO,      /* LSM.ref (soil reflectance, models.py:1917) */
    PROSAIL_O_LEAF_KA,  /* PROSPECT.ka = 1 - ks - kt (models.py:1066); prosail_prospect_* only */
    PROSAIL_O_LEAF_KE,  /* PROSPECT.ke = ks + ka     (models.py:1067) */
    PROSAIL_O_LEAF_OM,  /* PROSPECT.om = ks / ke     (models.py:1068) */
    PROSAIL_O_LAYER_R,  /* PROSPECT.r, .t, .Ra, .Ta, .denom: the one-plate terms the reference keeps as attributes */
    PROSAIL_O_LAYER_T,  /*   (models.py:959, 1019-1025); prosail_prospect_* only */
    PROSAIL_O_LAYER_RA,
    PROSAIL_O_LAYER_TA,
    PROSAIL_O_LAYER_DENOM,
    PROSAIL_NOUT
};

typedef struct prosail_handle prosail_handle_t;

/* PROSPECT(N, Cab, Cxc, Cbr, Cw, Cm, Can) parameter arrays, models.py:900.  Can may be NULL for version '5'. */
typedef struct prosail_leaf {
    const double *N, *Cab, *Cxc, *Cbr, *Cw, *Cm, *Can; /* [n] each */
} prosail_leaf_t;

/* LSM(reflectance, moisture), models.py:1908 */
typedef struct prosail_soil {
    const double *reflectance, *moisture; /* [n] each */
} prosail_soil_t;

/* SAIL(iza, vza, raa, ..., lai, hotspot, lidf_type, a, b), models.py:528-529 */
typedef struct prosail_canopy {
    int lidf_type;                               /* PROSAIL_LIDF_* */
    const double *lidf_a, *lidf_b;               /* [n]; lidf_b may be NULL for campbell */
    const double *lai, *hotspot;                 /* [n] */
    int n_geom;                                  /* geometries per parameter set (G >= 1) */
    int geom_per_sample;                         /* 0: iza/vza/raa are [n_geom], shared; 1: [n], n_geom must be 1 */
    const double *iza, *vza, *raa;               /* radians */
} prosail_canopy_t;

/* results; item index = sample * n_geom + geometry */
typedef struct prosail_out {
    double* plane[PROSAIL_NOUT]; /* spectra [n_items][pitch]; NULL = not wanted */
    long long pitch;
    unsigned mean_mask;          /* planes whose window means are wanted (bit = plane index) */
    double* means;               /* [n_items][popcount(mean_mask)][n_windows], planes in bit order */
    double* geom;                /* optional [n_items][PROSAIL_NGEOM]: VolScatt ks, ko, bf, Fs, Ft; tss, too, tsstoo (models.py:664-
                                    702); chi_s, chi_o, frho, ftau of the last leaf-inclination bin (what models.py:159 leaves behind) */
    int windows_only;            /* 1: evaluate only the wavelengths inside some window (planes must be NULL) */
} prosail_out_t;

/* fp32 mode (BASELINE north star: <= 1e-5 absolute against the float64 reference): spectra, planes and means are
 * float arrays; the per-set parameters stay float64 (the per-set prologue -- LIDF, volume scattering, hotspot -- is
 * always evaluated in float64, it is < 3% of the work and has real cancellations). */
typedef struct prosail_out_f32 {
    float* plane[PROSAIL_NOUT];
    long long pitch;
    unsigned mean_mask;
    float* means;
    double* geom;
    int windows_only;
} prosail_out_f32_t;

/* ---- lifetime --------------------------------------------------------------------------------------- */
/* Replaces the import-time global `lib = get_data_*()` (models.py:16-19): create a handle on `device`. */
int prosail_create(int device, prosail_handle_t** out);
void prosail_destroy(prosail_handle_t* h);
const char* prosail_last_error(void);
const char* prosail_version(void);

/* Upload one coefficient set.  Replaces PROSPECT.__set_coef (models.py:923-947) + library.py:43-64.
 * K* are the float32 tables [2101]; Kan may be NULL (version '5': zeros, models.py:935);
 * talf/t12/t21 are the per-band interface transmissivities in float64, computed by the host exactly as
 * PROSPECT.__calctav / __refl_trans_one_layer do (models.py:961-998, 1011-1015) for the chosen alpha. */
int prosail_set_leaf_tables(prosail_handle_t* h, int version, const float* Kab, const float* Kxc, const float* Kbr,
                            const float* Kw, const float* Km, const float* Kan, const double* talf, const double* t12,
                            const double* t21);
/* Soil spectra rsoil1/rsoil2 (library.py:57), float32 [2101]. */
int prosail_set_soil_tables(prosail_handle_t* h, const float* rsoil1, const float* rsoil2);
/* Inclusive wavelength windows [lo_nm, hi_nm] for the band means.  Replaces the hard-coded L8/ASTER windows of
 * SAIL.__store_L8/__store_aster (models.py:801-809, 836-841) and PROSPECT/LSM.select (models.py:1197-1225). */
int prosail_set_windows(prosail_handle_t* h, int n_windows, const int* lo_nm, const int* hi_nm);
/* The same windows with a spectral response: window w carries one weight per wavelength lo_nm[w]..hi_nm[w] (1 nm steps;
 * the arrays of all windows concatenated in `weights`), and its "mean" becomes sum(w x) / sum(w) -- a true sensor-band
 * convolution.  The reference only has the unweighted special case (all weights 1, models.py:811-851); SURVEY.md 8(f)-2. */
int prosail_set_windows_srf(prosail_handle_t* h, int n_windows, const int* lo_nm, const int* hi_nm, const double* weights);

/* ---- the hot path ------------------------------------------------------------------------------------ */
/* pyrism.PROSPECT(...) for n parameter sets (models.py:900-921).  out->plane[]: PROSAIL_O_LEAF_R (ks), _LEAF_T (kt), _LEAF_KA/KE/OM
 * (models.py:1064-1068) and _LAYER_R/T/RA/TA/DENOM (the one-plate attributes, models.py:959), each [n][out->pitch], NULL = not
 * wanted; the other fields of `out` must be zero.
 * means_f32 (optional) = float32 window means of (ks, kt, ka, ke, om), [n][5][n_windows]  (models.py:1071-1195). */
int prosail_prospect_f64(prosail_handle_t* h, int version, long long n, const prosail_leaf_t* leaf, const prosail_out_t* out,
                         float* means_f32, int on_host, void* stream);

/* pyrism.LSM(reflectance, moisture) for n parameter sets (models.py:1908-1920): rho [n][pitch];
 * means_f32 (optional) [n][1][n_windows] float32 (models.py:1918-2006). */
int prosail_soil_f64(prosail_handle_t* h, long long n, const prosail_soil_t* soil, double* rho, long long pitch,
                     float* means_f32, int on_host, void* stream);

/* pyrism.VolScatt(iza, vza, raa).coef(lidf_type, a, b) + LIDF.campbell/verhoef (models.py:82-175, 283-392),
 * plus the direct transmittances/hotspot terms of SAIL.__calc (models.py:664-702) when lai/hotspot are given.
 * geom [n_items][PROSAIL_NGEOM] (see prosail_out_t.geom); lidf (optional) [n][18]. */
int prosail_volscatt_f64(prosail_handle_t* h, long long n, const prosail_canopy_t* canopy, double* geom, double* lidf,
                         int on_host, void* stream);

/* LIDF.campbell(a, n_elements) / LIDF.verhoef(a, b, n_elements) and VolScatt(...).coef(lidf_type, n_elements, a, b) for a number
 * of leaf-inclination classes other than SAIL's 18 (the reference's n_elements argument, models.py:82-175, 283-392).
 * lidf_a, lidf_b [n] (lidf_b may be NULL for Campbell); iza, vza, raa [n_geom] radians, normalised (unused when geom is NULL);
 * lidf [n][n_elements] is always written; geom (optional) [n][n_geom][PROSAIL_NGEOM]: ks, ko, bf, Fs, Ft in slots 0-4,
 * chi_s, chi_o, frho, ftau of the last class in slots 8-11, slots 5-7 zero.  1 <= n_elements <= 4096. */
int prosail_volscatt_bins_f64(prosail_handle_t* h, long long n, int lidf_type, const double* lidf_a, const double* lidf_b, int n_geom,
                              const double* iza, const double* vza, const double* raa, int n_elements, double* geom, double* lidf,
                              int on_host, void* stream);

/* pyrism.VolScatt(iza, vza, raa).volume(lza) (models.py:177-258): interception and bidirectional scattering terms of ONE leaf
 * zenith angle lza_deg (degrees) for n_geom geometries (radians, normalised); out [n_geom][4] = chi_s, chi_o, frho, ftau. */
int prosail_volume_f64(prosail_handle_t* h, int n_geom, const double* iza, const double* vza, const double* raa, double lza_deg,
                       double* out, int on_host, void* stream);

/* pyrism.SAIL(...) for n parameter sets x n_geom geometries (models.py:528-729) with leaf and soil spectra
 * supplied as arrays: ks, kt, rho are [n][pitch_*] or a single spectrum broadcast to all sets (pitch_* = 0). */
int prosail_sail_f64(prosail_handle_t* h, long long n, const double* ks, long long pitch_ks, const double* kt,
                     long long pitch_kt, const double* rho, long long pitch_rho, const prosail_canopy_t* canopy,
                     const prosail_out_t* out, int on_host, void* stream);

/* Fused PROSPECT -> LSM -> SAIL (examples/prospect_sail.py:9-28 per parameter set) without materialising the leaf
 * and soil spectra in HBM: the LUT-production kernel. */
int prosail_prosail_f64(prosail_handle_t* h, int version, long long n, const prosail_leaf_t* leaf, const prosail_soil_t* soil,
                        const prosail_canopy_t* canopy, const prosail_out_t* out, int on_host, void* stream);

/* fp32 mode of the four entry points above: same semantics, float spectra in and out (tolerance 1e-5 absolute). */
int prosail_prospect_f32(prosail_handle_t* h, int version, long long n, const prosail_leaf_t* leaf, const prosail_out_f32_t* out,
                         float* means_f32, int on_host, void* stream);
int prosail_soil_f32(prosail_handle_t* h, long long n, const prosail_soil_t* soil, float* rho, long long pitch,
                     float* means_f32, int on_host, void* stream);
int prosail_sail_f32(prosail_handle_t* h, long long n, const float* ks, long long pitch_ks, const float* kt,
                     long long pitch_kt, const float* rho, long long pitch_rho, const prosail_canopy_t* canopy,
                     const prosail_out_f32_t* out, int on_host, void* stream);
int prosail_prosail_f32(prosail_handle_t* h, int version, long long n, const prosail_leaf_t* leaf, const prosail_soil_t* soil,
                        const prosail_canopy_t* canopy, const prosail_out_f32_t* out, int on_host, void* stream);

/* ---- LUT inversion (the consumer of the LUT; BASELINE north star; no counterpart in the reference) ------
 * For each of n_obs observed feature vectors (e.g. the 15 window means of a pixel) the index of the LUT entry with the smallest
 *   cost(q, e) = sum_d ( s_d obs[q][d] - s_d lut[e][d] )^2,   s_d = sqrtf(weights[d])  (1 when weights is NULL),
 * evaluated in float32 in ascending d with fmaf(diff, diff, acc) on pre-scaled features; ties go to the lowest index;
 * best_index = local index + index_offset (the first LUT row this shard holds; -1 if every cost is NaN/inf).
 * n_feat <= 32; lut [n_lut][pitch_lut], obs [n_obs][pitch_obs] are float32 (the fp32-mode means of prosail_prosail_f32, or any
 * other feature table); weights is a HOST array of n_feat values in both modes.  Multi-GPU: every rank inverts against its own
 * LUT shard and the per-rank (cost, index) pairs are reduced by the caller (pyrism_b200.inversion does it over NCCL). */
int prosail_invert_f32(prosail_handle_t* h, long long n_lut, int n_feat, const float* lut, long long pitch_lut, long long n_obs,
                       const float* obs, long long pitch_obs, const float* weights, long long index_offset, long long* best_index,
                       float* best_cost, int on_host, void* stream);

/* How the last prosail_invert_f32 call on this handle ran (waits for it): used_tensor_cores = 1 when the candidate pass ran on the
 * tcgen05 tensor cores (n_feat <= 15, n_lut >= 512; csrc/invert_tc.cuh) -- the result is the same bit-exact contract either way;
 * redone_exactly = observations whose candidate list could not be proven complete and were rescanned exactly; timed_out != 0 if a
 * bounded device-side wait expired (the observations of that CTA were rescanned exactly as well).  The error model of the candidate
 * pass is monitored in every call: error_ratio = largest measured |approximate - real| / sum_k |A_k||B_k| over all kept candidates
 * (to be compared with the model's KAPPA = 8e-6), bound_violations = candidates found outside the model's bound (their
 * observations were rescanned exactly; expected 0).  Any pointer may be NULL.  No counterpart in the reference. */
int prosail_invert_last_stats(prosail_handle_t* h, int* used_tensor_cores, long long* redone_exactly, int* timed_out, double* error_ratio,
                              long long* bound_violations);

/* The k best entries per observation (1 <= k <= 16), for estimators that average the parameters of the k closest entries:
 * best_index / best_cost are [n_obs][k], ascending cost, equal costs by ascending index; -1 / +inf pad when fewer than k
 * entries have a finite cost.  Same cost definition and evaluation order as prosail_invert_f32 (k = 1 gives the same result). */
int prosail_invert_topk_f32(prosail_handle_t* h, int k, long long n_lut, int n_feat, const float* lut, long long pitch_lut,
                            long long n_obs, const float* obs, long long pitch_obs, const float* weights, long long index_offset,
                            long long* best_index, float* best_cost, int on_host, void* stream);

/* ---- memory / stream helpers (so that a host without torch/cupy can drive the device API) -------------- */
int prosail_dev_alloc(prosail_handle_t* h, long long bytes, void** out);
int prosail_dev_free(prosail_handle_t* h, void* p);
int prosail_host_alloc(long long bytes, void** out); /* pinned */
int prosail_host_free(void* p);
int prosail_memcpy_h2d(prosail_handle_t* h, void* dst, const void* src, long long bytes, void* stream);
int prosail_memcpy_d2h(prosail_handle_t* h, void* dst, const void* src, long long bytes, void* stream);
int prosail_sync(prosail_handle_t* h, void* stream);
/* CUDA-event timer on `stream` (NULL = handle stream): begin/end bracket device work, end returns milliseconds. */
int prosail_timer_begin(prosail_handle_t* h, void* stream);
int prosail_timer_end(prosail_handle_t* h, void* stream, float* ms);
/* number of kernel launches issued through this handle so far */
long long prosail_launch_count(prosail_handle_t* h);
/* The LIDF / volume-scattering / hotspot prologue exists in two mappings: one warp per (set, geometry) with the 18 leaf
 * bins across lanes and shuffle reductions (lowest latency, small batches) and one thread per (set, geometry)
 * (highest throughput).  Batches of at least `min_items` items use the latter (default 2048; 0 = always). */
int prosail_set_prologue_threshold(prosail_handle_t* h, long long min_items);
/* Per-kernel device timing: while enabled, every launch is bracketed by CUDA events on its stream.
 * prosail_profile_read synchronises, returns summed milliseconds and launch counts for
 * [0] the band kernel (PROSPECT/4SAIL arithmetic), [1] the prologue, [2] the band-mean finalisation, and resets. */
int prosail_profile_enable(prosail_handle_t* h, int on);
int prosail_profile_read(prosail_handle_t* h, double ms[3], long long count[3]);
/* name, SM count, HBM bytes of the handle's device */
int prosail_device_info(prosail_handle_t* h, char* name, int name_len, int* sm_count, long long* mem_bytes);

/* PCI bus id of the handle's device ("0000:1b:00.0"), so that a host process can bind itself and its pinned buffers to
 * the NUMA node the GPU hangs off before it allocates them (pyrism_b200/engine.py does). */
int prosail_device_pci_bus_id(prosail_handle_t* h, char* out, int len);

/* ---- calibration / self-test -------------------------------------------------------------------------- */
/* FP64 roofline denominator: sustained DFMA throughput of the device (TFLOP/s, FMA = 2 flops), measured with a
 * register-resident FMA kernel for about `millis` milliseconds. */
int prosail_measure_fp64_peak(prosail_handle_t* h, int millis, double* tflops);
/* evaluates one fp64 primitive of the kernels on device arrays (op: 0 rcp, 1 sqrt, 2 exp(x<=0), 3 log2, 4 exp2,
 * 5 tau(k) = (1-k)e^-k + k^2 E1(k)); used by the GPU unit tests. */
int prosail_test_math(prosail_handle_t* h, int op, long long n, const double* in, double* out, int on_host);

#ifdef __cplusplus
}
#endif
#endif /* PROSAIL_B200_H */
