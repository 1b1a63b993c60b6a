// sm_100a kernels for the batched PROSPECT-5/D -> 4SAIL -> band-mean path.
//
// Mapping (band kernel): one THREAD owns NU wavelengths (b, b + 32, ...; NU = 2 or 3, struct Cfg); its once-per-item constants
// (absorption coefficients, soil spectra and -- fused fp64 instances -- the folded interface terms) live in shared memory as
// per-thread runs of plane pairs (Cfg::koff), the other instances keep the interface terms in registers;
// a WARP owns NU x 32 consecutive bands (66 32-band tiles cover 2101 bands, 99.5% lane efficiency) and walks
// over the work items (parameter set x geometry).  The NU bands of a thread are independent
// dependency chains in one instruction stream: the path is bound by the FP64 pipe (DFMA latency 8.6
// cycles at one warp instruction per 2 cycles per SM sub-partition, measured), and with 3-4 resident warps
// per sub-partition it is this instruction-level parallelism that keeps the pipe fed.  Rare cases (arguments outside the tau
// table, zero absorption, the J1 series branch) are flagged in the branch-free fast path and repaired
// afterwards, so that the hot loop is a handful of large basic blocks the scheduler can interleave.
// The kernel is a template over the real type: double is the parity path (<= 1e-9 against the reference), float the
// fp32 mode (<= 1e-5), whose item loop runs in packed FP32 (FFMA2 / FMUL2 / FADD2: the two bands of a thread per instruction).
// Per-item scalars come from 128-byte records that the prologue kernel packs and that each CTA stages
// into shared memory with bulk async copies (cp.async.bulk + mbarrier, double buffered), so the inner
// loop touches HBM only for the coalesced 256-byte-per-warp stores of the output spectra.
//
// Reference (ibaris/pyrism 0.0.5, pyrism/models/models.py) lines are cited at each step.
#pragma once
#include <type_traits>

#include "math32.cuh"

namespace pb {

#ifndef PB_MIN_CTAS
#define PB_MIN_CTAS 1   /* resident CTAs per SM the band kernel is compiled for (16 warps x 128 registers fill the register file) */
#endif

#ifndef PB_CHUNK
#define PB_CHUNK 8
#endif
#ifndef PB_WARPS
#define PB_WARPS 16     /* warps per CTA (one CTA per SM); warps are independent (own staging buffers), the CTA only shares the math tables */
#endif
// fp32 mode: warps per CTA and resident CTAs per SM.  Round 2n: ONE CTA of 24 warps (80 registers per thread: no spills in the item
// loop) instead of two of 16 (64 registers, LDL 2.6 / STL 2.0 per item): BRF 7.47e7 -> 7.96e7 spectra/s, band means 1.36e8 -> 1.37e8;
// 20 warps x 96 registers: 7.89e7 / 1.28e8, 2 x 12 warps: 7.48e7 / 1.29e8 (profiles/r2n_variants_b.txt)
#ifndef PB_F32_WARPS
#define PB_F32_WARPS 24
#endif
#ifndef PB_F32_CTAS
#define PB_F32_CTAS 1
#endif

constexpr int NB = 2101;          // wavelengths 400..2500 nm
constexpr int NTILES = 66;        // 32-band tiles (unit of the band-mean window map)
constexpr int NBP = NTILES * 32;  // padded band count (2112)
constexpr int NT2 = NTILES / 2;   // most tile groups a kernel configuration can have (NU = 2: 33 groups of 64 bands)

// Kernel configuration: NU wavelengths per thread (NU independent dependency chains in one instruction stream) and
// WARPS warps per CTA (one CTA per SM in fp64; the register file holds WARPS x 32 x (65536 / (WARPS x 32)) registers).
//   Cfg<2, 16>: 128 registers per thread, 4 warps x 2 chains per SM sub-partition
//   Cfg<3, 12>: 168 registers per thread, 3 warps x 3 chains per SM sub-partition (22 groups of 96 bands)
template <int NU_, int WARPS_>
struct Cfg {
    static constexpr int NU = NU_;                     // bands per thread: band0 + 32 u
    static constexpr int WARPS = WARPS_;
    static constexpr int THREADS = WARPS_ * 32;
    // SailSmem::kc is [plane pair][thread][plane of the pair][u]: the 2 NU values a thread holds of two consecutive planes are one
    // contiguous, 16-byte aligned run (NU = 3: 48 bytes = three conflict-free 128-bit loads instead of six 64-bit ones -- a thread's
    // stride of 48 bytes puts the 8 lanes of a quarter warp on 8 different 16-byte bank groups)
    static constexpr int KPAIR = 2 * NU_ * THREADS;    // elements per plane pair
    static constexpr int KCOL = 2 * NU_;               // elements per thread and pair
    __host__ __device__ static constexpr int koff(int plane) { return (plane >> 1) * KPAIR + (plane & 1) * NU_; }   // offset of (plane, u = 0) from the thread's base
    static constexpr int NGROUPS = NTILES / NU_;       // groups of NU 32-band tiles, one per warp
    static_assert(NTILES % NU_ == 0, "the 66 tiles must split evenly into groups");
};
constexpr int CHUNK = PB_CHUNK;   // work items per shared-memory stage (per warp)
constexpr int REC = 16;           // values per record (128 B in double, 64 B in float)
constexpr int MAX_TILE_SLOTS = 8; // windows that may overlap one 32-band tile
constexpr int MAX_WINDOWS = 64;
constexpr int MAX_SLOTS = 256;

// planes of the per-band constant array  T[NCONST][NBP]
enum { C_KAB = 0, C_KXC, C_KAN, C_KBR, C_KW, C_KM, C_TALF, C_T12, C_T21, C_RS1, C_RS2, C_RALF, C_R12, C_R21, NCONST };
// (C_RALF.. = 1 - talf, 1 - t12, 1 - t21: only the float image carries them, rounded once from the float64 values)

// sample record (per parameter set)
enum { S_CAB = 0, S_CXC, S_CAN, S_CBR, S_CW, S_CM, S_NM1, S_SOIL1, S_SOIL2, S_DDB, S_DDF, S_NEGLAI, S_LAI, S_LMAX, S_THR, S_BF };
// In records WITH a leaf (fused PROSAIL, leaf-only) the slots S_DDB, S_DDF are not needed by the kernels (the P / Q form of
// sail_diffuse) and carry S_CWX, S_CBRX instead: copies of Cw/N and Cbr/N that are NaN when a pigment content whose term a tile group
// skips is not finite (leaf_kall_all).
enum { S_CWX = S_DDB, S_CBRX = S_DDF };
#if !PB_RINF_P || !PB_F32_ALG
#error "S_CWX / S_CBRX reuse the slots of ddb / ddf: the fused kernels must use the P / Q form of sail_diffuse (PB_RINF_P = PB_F32_ALG = 1)"
#endif
// (S_NEGLAI: -lai, +0.0 for bare soil (the no-canopy flag); double records with PB_EXPS: -lai 2048 log2(e).  S_LMAX: 500/|N-1|, the
//  clamp of log2 b in pow_pos_scaled.)
// geometry record (per parameter set x geometry)
enum { G_K = 0, G_KO, G_SDB, G_SDF, G_DOB, G_DOF, G_FS, G_FT, G_TSS, G_TOO, G_TSSTOO, G_LAISUM, G_Z, G_FSL, G_FTL, G_DTT };
// (G_FSL, G_FTL: Fs lai sumint, Ft lai sumint -- the single-scattering term rsos = w lai sumint of models.py:706-710 as two FMAs on rho, tau;
//  G_DTT: tsstoo - tss too, the hotspot excess of the bidirectional gap fraction -- see the soil coupling in sail_couple)

// output planes (bit index in out_mask / mean_mask)
enum { O_RSOT = 0, O_RDDT, O_RSDT, O_RDOT, O_RSO, O_RDD, O_TDD, O_RSD, O_TSD, O_RDO, O_TDO, O_KE, O_LEAF_R, O_LEAF_T, O_RHO,
       O_LEAF_KA, O_LEAF_KE, O_LEAF_OM,           // PROSPECT's ka, ke, om (models.py:1066-1068), PROSPECT mode only
       O_LAYER_R, O_LAYER_T, O_LAYER_RA, O_LAYER_TA, O_LAYER_DENOM,   // one-plate r, t, Ra, Ta, denom (models.py:959, 1019-1025), idem
       NOUT };
constexpr int NGEOM = 12;         // per-item scalars returned to the host: ks ko bf Fs Ft tss too tsstoo + chi_s chi_o frho ftau (last leaf bin)

struct WindowMap {                       // which band-mean windows touch which warp tile
    int nslots[NTILES];
    unsigned mask[NTILES][MAX_TILE_SLOTS];   // lanes of the tile inside the window
    int slot[NTILES][MAX_TILE_SLOTS];        // global slot id (window-major)
    int wstart[MAX_WINDOWS + 1];             // slots of window w: [wstart[w], wstart[w+1])
    int wcount[MAX_WINDOWS];                 // number of bands in window w
    double wsum[MAX_WINDOWS];                // sum of the spectral-response weights of window w (= wcount when unweighted)
    int nwindows, ntotal;
    int ntiles_active;                       // 32-band tiles touched by at least one window
    int active_tile[NTILES];
    int ngroups_active[2];                   // [NU - 2]: groups of NU tiles touched by at least one window
    int active_group[2][NT2];
    // Packed variant of the map (band-mean-only LUTs, plain boxcar windows): the wavelengths that lie inside at least one
    // window, in ascending order, are the ONLY lanes that exist -- 782 of 2101 for the reference's 6 L8 + 9 ASTER windows
    // (25 tiles instead of the 38 tiles = 19 two-tile groups the unpacked map touches).  A window is still a run of
    // consecutive lanes (the list is sorted), so the segmented warp reduction is unchanged; `band` maps a packed lane back to
    // its wavelength index for the gather of the per-band constants (lanes past `nbands_packed`: band 0, in no mask).
    // In the packed map the slots are PIECES: maximal runs of consecutive lanes of one tile that belong to the same set of
    // windows (cut at every window edge and every tile edge).  Pieces are disjoint, so ONE segmented warp scan per tile yields
    // all of its piece sums (5 shuffle + add steps, however many windows overlap there), and a window is the run of pieces
    // [wstart[w], wend[w]) -- runs of different windows overlap where the windows do.
    int packed;                              // 1: tiles / masks / slots of this map are in packed-lane space
    int nbands_packed;
    int wend[MAX_WINDOWS];                   // one past the last slot of window w (unpacked map: wstart[w + 1])
    int band[NBP];
    int piece[NBP];                          // packed map: slot (piece) id of every lane, -1 for the padding lanes
};

template <typename T>
struct SailArgs {
    const T* bandc;            // [NCONST][NBP]
    const void* tables;        // Tables<T>::type image in global memory
    const WindowMap* win;      // band-mean window map (global memory, read uniformly)
    const double* srf;         // [slots][32] spectral-response weight of each lane of each window segment; NULL = boxcar (the reference)
    const T* srec;             // [nsamples][REC]
    const T* grec;             // [nitems][REC]
    long long nitems;
    int G;                     // geometries per sample (item = s*G + g)
    int ngroups;               // tile groups visited (all NTILES / NU, or the active ones in windows-only mode)
    int nwarps_main;           // R x ngroups warps that own one tile group each and stride over the chunks below nchunks_main (see k_bands)
    int nwarps;                // warps that work; those past nwarps_main share the (group, chunk) pairs of the chunks from nchunks_main on
    long long nchunks_main;
    // SAIL-only mode: leaf / soil spectra from memory, row pitch in elements (0 = one spectrum for all samples)
    const T* in_r; const T* in_t; const T* in_rho;
    long long in_pitch_r, in_pitch_t, in_pitch_rho;
    T* out[NOUT];              // output planes [nitems][pitch] (NULL = not requested)
    long long out_pitch;
    unsigned mean_mask;        // planes whose band means are accumulated
    T* partial;                // [nitems][nmean][ntotal slots]
    int nmean;
    int only_active_tiles;     // 1: visit only tiles touched by a window (band-mean-only LUTs)
    int packed;                // 1: `win` is the packed map (lanes = in-window wavelengths only); implies only_active_tiles, no planes
    int f32_means;             // 1: round values to float before averaging (PROSPECT/LSM convention, models.py:1071)
};


// ---------------------------------------------------------------------------------------------
// bulk-async staging helpers (mbarrier + cp.async.bulk, CTA scope)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n"
        " bra WAIT_%=;\n"
        "DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T>
struct alignas(128) Stage {
    T s[CHUNK * REC];        // sample records spanned by the chunk
    T g[CHUNK * REC];        // geometry records of the chunk
};

// once-per-item band constants live in shared memory, not registers: kab kxc kan kbr kw km rs1 rs2, and in the double kernels
// also the five (folded) interface terms.  Every one of them is read once per item, and 20 registers less per thread are what
// keeps the fp64 item loop free of local-memory spills (with 200+ KB of the SM's L1 configured as shared memory, spill
// reloads miss L1 and show up as long-scoreboard stalls: 4-12% of the warp time in the ncu captures of the spilling builds).
template <typename T, int MODE> __host__ __device__ constexpr bool consts_in_smem() { return !std::is_same<T, float>::value && MODE == 0 /* MODE_PROSAIL */; }
template <typename T, int MODE> __host__ __device__ constexpr int kplanes() { return consts_in_smem<T, MODE>() ? 13 : 8; }

// (A fully thread-major layout [thread][plane][u] measured slower in round 2f: BRF -0.8 %, leaf-only -11 %, profiles/r2f_variants.txt.)
template <typename T, typename CF, int MODE>
struct alignas(128) SailSmem {
    typename Tables<T>::type mt;           // shared by the CTA, read-only after the prologue
    Stage<T> stage[CF::WARPS][2];          // per-warp double-buffered record staging: no CTA-wide barrier in the loop
    T kc[((kplanes<T, MODE>() + 1) / 2) * CF::KPAIR];   // [plane pair][thread][plane of the pair][u], see Cfg::koff
    uint64_t bar[CF::WARPS][2];
};

// issue the two bulk copies of one chunk (called by one thread)
template <typename T>
__device__ __forceinline__ void stage_issue(const SailArgs<T>& a, Stage<T>* st, uint64_t* bar, long long chunk) {
    const long long i0 = chunk * CHUNK;
    const long long i1 = min(i0 + (long long)CHUNK, a.nitems);
    // item indices of one launch fit in 32 bits (run_device_t slabs: <= 2^20 items): no 64-bit division in the loop
    const long long s0 = (unsigned)i0 / (unsigned)a.G, s1 = (unsigned)(i1 - 1) / (unsigned)a.G;
    const uint32_t gbytes = (uint32_t)(i1 - i0) * REC * (uint32_t)sizeof(T);
    const uint32_t sbytes = (uint32_t)(s1 - s0 + 1) * REC * (uint32_t)sizeof(T);
    mbar_expect_tx(bar, gbytes + sbytes);
    bulk_g2s(st->g, a.grec + i0 * REC, gbytes, bar);
    bulk_g2s(st->s, a.srec + s0 * REC, sbytes, bar);
}

// ---------------------------------------------------------------------------------------------
// per-band device functions (T = double: the parity path; T = float: the fp32 mode, <= 1e-5 absolute)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct BandConst {
    T talf, ralf, t12, r12, t21, r21;  // interface terms (models.py:1011-1016)
    T ta21, t1221;                     // talf t21, t12 t21: the folded factors of the double fast path (leaf_layers_fast)
};
enum { KC_KAB = 0, KC_KXC, KC_KAN, KC_KBR, KC_KW, KC_KM, KC_RS1, KC_RS2,     // planes of SailSmem::kc
       KC_TA21, KC_RALF, KC_T1221, KC_R12, KC_R21 };                          // double only: talf t21, 1 - talf, t12 t21, 1 - t12, 1 - t21

// interface constants of one band from the constant array (bc = bandc + band)
template <typename T>
__device__ __forceinline__ BandConst<T> load_band_const(const T* __restrict__ bc) {
    BandConst<T> c;
    c.talf = bc[C_TALF * NBP]; c.t12 = bc[C_T12 * NBP]; c.t21 = bc[C_T21 * NBP];
    if constexpr (std::is_same<T, float>::value) {      // the complements are uploaded (rounded from float64) instead of being recomputed in float
        c.ralf = bc[C_RALF * NBP]; c.r12 = bc[C_R12 * NBP]; c.r21 = bc[C_R21 * NBP];
    } else {
        c.ralf = T(1) - c.talf; c.r12 = T(1) - c.t12; c.r21 = T(1) - c.t21;                          // models.py:1012-1016
    }
    c.ta21 = c.talf * c.t21; c.t1221 = c.t12 * c.t21;
    return c;
}

template <typename T> __device__ __forceinline__ constexpr bool is_f32() { return std::is_same<T, float>::value; }

// PROSPECT for one band and one parameter set: leaf reflectance (ks) and transmittance (kt).
// models.py:949-959 (kall, tau), 1019-1025 (one layer), 1043-1068 (Stokes, N layers).
template <typename CF, typename T>
__device__ __forceinline__ T leaf_kall(const T* __restrict__ kc, const T* __restrict__ s) {
    // kall = sum_i C_i K_i / N ; the division by N is folded into the record (C_i / N)      :950-951
    T kall = s[S_CAB] * kc[CF::koff(KC_KAB)];
    kall = fma(s[S_CXC], kc[CF::koff(KC_KXC)], kall);
    // PROSPECT-5 has no anthocyanin term (Kan = zeros, models.py:935): the record carries Can/N = +0.0 and the FMA is
    // predicated off by an integer test on that (warp-uniform) value -- one three-register DFMA less per band and item
    if (!is_zero(s[S_CAN])) kall = fma(s[S_CAN], kc[CF::koff(KC_KAN)], kall);
    kall = fma(s[S_CBR], kc[CF::koff(KC_KBR)], kall);
    kall = fma(s[S_CW], kc[CF::koff(KC_KW)], kall);
    return fma(s[S_CM], kc[CF::koff(KC_KM)], kall);
}
// The 2 NU values a thread holds of the plane pair starting at the (even) plane `first`, with explicit 128-bit loads (the run is 16-byte
// aligned, see Cfg::koff): a[u] = plane `first`, b[u] = plane `first + 1`.
template <typename CF, typename T>
__device__ __forceinline__ void load_pair(const T* __restrict__ kc, int first, T (&a)[CF::NU], T (&b)[CF::NU]) {
    constexpr int NU = CF::NU;
    T f[2 * NU];
    const T* p = kc + CF::koff(first);
    if constexpr (std::is_same<T, float>::value && NU == 2) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
    } else if constexpr (!std::is_same<T, float>::value) {
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const double2 q = reinterpret_cast<const double2*>(p)[i];
            f[2 * i] = q.x; f[2 * i + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 2 * NU; ++i) f[i] = p[i];
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) { a[u] = f[u]; b[u] = f[NU + u]; }
}
// kall of all NU bands of a thread (:950-951), the six absorption tables read as three plane pairs
// CAN = false: instances launched for PROSPECT-5 batches only (dispatch_bands), where the anthocyanin term does not exist
// (Kan = zeros, models.py:935): not even a predicated-off FMA is issued for it
// kz (warp-uniform, set per tile group by k_bands): which pigment tables are identically zero over the group's bands --
//   0: none;  1: Kab and Kxc (every band from 784 nm on: chlorophyll ends at 779-780 nm, carotenoids at 559-560 nm);
//   2: Kab, Kxc, Kan and Kbr (every band from 1100 nm on: 14 of the 22 96-band groups), where kall = Cw Kw + Cm Km.
// The skipped terms are exact zeros (a finite content times 0.0), so the result is bit for bit that of the full sum.  A non-finite
// content is the exception -- NaN or inf times 0.0 is NaN, which the reference turns into tau = 1 (models.py:953: `kall > 0` is
// False) -- so the paths that skip read the first content they do use from a copy that the prologue sets to NaN when one of the
// skipped contents is not finite (S_CBRX for kz = 1, S_CWX for kz = 2): kall is NaN exactly where the reference's is
// (tests/test_gpu_parity.py::test_non_finite_pigment_content...).  Per band and item: 2.7 instead of 5 FP64 instructions on average.
template <typename CF, bool CAN, typename T>
__device__ __forceinline__ void leaf_kall_all(const T* __restrict__ kc, const T* __restrict__ s, int kz, T (&kall)[CF::NU]) {
    constexpr int NU = CF::NU;
    static_assert(KC_KAB == 0 && KC_KXC == 1 && KC_KAN == 2 && KC_KBR == 3 && KC_KW == 4 && KC_KM == 5, "plane pairs");
    T a[NU], b[NU];
    if (kz == 2) {
        load_pair<CF>(kc, KC_KW, a, b);
#pragma unroll
        for (int u = 0; u < NU; ++u) kall[u] = fma(s[S_CM], b[u], s[S_CWX] * a[u]);
        return;
    }
    if (kz == 1) {
        load_pair<CF>(kc, KC_KAN, a, b);
#pragma unroll
        for (int u = 0; u < NU; ++u) kall[u] = s[S_CBRX] * b[u];
    } else {
        load_pair<CF>(kc, KC_KAB, a, b);
#pragma unroll
        for (int u = 0; u < NU; ++u) kall[u] = fma(s[S_CXC], b[u], s[S_CAB] * a[u]);
        load_pair<CF>(kc, KC_KAN, a, b);
#pragma unroll
        for (int u = 0; u < NU; ++u) kall[u] = fma(s[S_CBR], b[u], kall[u]);
    }
    // PROSPECT-5 has no anthocyanin term (Kan = zeros, models.py:935): CAN = false builds; in the PROSPECT-D builds the record of a
    // '5' call still carries Can/N = +0.0 and the FMA is skipped on that (warp-uniform) value
    if constexpr (CAN) {
        if (!is_zero(s[S_CAN])) {
#pragma unroll
            for (int u = 0; u < NU; ++u) kall[u] = fma(s[S_CAN], a[u], kall[u]);
        }
    }
    load_pair<CF>(kc, KC_KW, a, b);
#pragma unroll
    for (int u = 0; u < NU; ++u) kall[u] = fma(s[S_CM], b[u], fma(s[S_CW], a[u], kall[u]));
}
template <typename CF, typename T>
__device__ __forceinline__ void soil_rho_all(const T* __restrict__ kc, const T* __restrict__ s, T (&rs)[CF::NU]) {
    static_assert(KC_RS1 == 6 && KC_RS2 == 7, "plane pairs");
    T a[CF::NU], b[CF::NU];
    load_pair<CF>(kc, KC_RS1, a, b);
#pragma unroll
    for (int u = 0; u < CF::NU; ++u) rs[u] = fma(s[S_SOIL1], a[u], s[S_SOIL2] * b[u]);
}
// soil reflectance of the LSM mix, models.py:1917
template <typename CF, typename T>
__device__ __forceinline__ T soil_rho(const T* __restrict__ kc, const T* __restrict__ s) {
    return fma(s[S_SOIL1], kc[CF::koff(KC_RS1)], s[S_SOIL2] * kc[CF::koff(KC_RS2)]);
}

// Branch-free plate + Stokes stage.  Returns true when 1 - r - t < 2^-17 (no or almost no absorption; r + t >= 1 is the
// zero-absorption branch of models.py:1057-1059); the values written are then not good enough -- b^(2(N-1)) - 1 and D are tiny
// and the second-order steps / economised polynomials of this path are amplified by 1/D -- and leaf_layers_slow() must be
// called for that band.  Realistic leaves (k >= 1e-4) never get there.
// The interface constants enter in FOLDED form (ta21 = talf t21, t1221 = t12 t21), so that Ta = ta21 tau/(1 - x^2) and
// t = t1221 tau/(1 - x^2) cost one multiplication each.
template <typename T, typename MT>
__device__ __forceinline__ bool leaf_layers_fast(T r21, T ta21, T t1221, T ralf, T r12, T tau, T nm1, int lmax_hi, const MT& mt, T& refl, T& tran) {
    const T one = T(1), half = T(0.5);
    // one layer                                                                              :1019-1025
    const T x = r21 * tau;
    const T tq = tau * rcp_b(fma(-x, x, one));     // benign: scales r - r12, t, Ta, Ra - ralf consistently (a slightly different plate)
    const T Ta = ta21 * tq;
    const T t = t1221 * tq;
    const T Ra = fma(x, Ta, ralf);
    const T r = fma(x, t, r12);
    // Stokes: D^2 = (1+r+t)(1+r-t)(1-r+t)(1-r-t)                                             :1043
    const T q = one - r;
    const T r2 = r * r;
    const T A = one + fma(-t, t, r2);              // 1 + r^2 - t^2  (> 0: r^2 - t^2 > -1)
    const T P12 = fma(T(2), r, A);                 // (1+r+t)(1+r-t) = (1+r)^2 - t^2: >= 1, no cancellation in this form
    const T qmt = q - t;                           // 1 - r - t: <= 0 flags zero absorption
    const T P34 = (q + t) * qmt;                   // the small factor keeps its product form
    const T D = sqrt_b_raw(P12 * P34);             // benign: D is itself the small quantity (a - 1/a = D/r).  P34 <= 0 only when r + t >= 1: that lane is flagged and recomputed
    const T hD = half_of(D);                       // D / 2 as an exponent decrement (integer pipe); D > 0 on every lane that is kept
    const T al = fma(half, A, hD);                 // a * r   (a of :1046)
    const T be = fma(-half, A, one + hD);          // b * t   (b of :1047), using 1 - r^2 + t^2 = 2 - A
    const T b = be * rcp_k(t);                     // NOT benign: b^(2(N-1)) - 1 differences it against 1 (third-order step)
    // b^(N-1) = 2^((N-1) log2 b)                                                             :1049
#if PB_POWS
    const T bNm1 = pow_pos_scaled(b, nm1, lmax_hi, mt);
#else
    const T bNm1 = pow_pos(b, nm1, mt);
#endif
    const T bN2 = bNm1 * bNm1;
    // With a = al/r:  kt = Ta bNm1 (al^2 - r^2)/X,  ks = Ra + Ta t al r (bN2 - 1)/X,
    //                 X = al^2 bN2 - r^2 - al r^2 (bN2 - 1)      (algebraically :1052-1065, one division)
    //                   = (al^2 - r^2) + (bN2 - 1) al (al - r^2)
    const T bm1 = bN2 - one;
#if PB_STOKES_Y
    // Round 2n: divide the X form below by a = al/r once more.  With 1/a = (A - D)/(2r) (because A^2 - D^2 = 4 r^2):
    //   Rsub = r bm1/(al bN2 - gl),  Tsub = bNm1 D/(al bN2 - gl),  gl = (A - D)/2,   and   1 - Rsub r = Y/(al bN2 - gl)  with
    //   Y = al bN2 - gl - r^2 bm1 = bm1 (al - r^2) + (al - gl) = bm1 (al - r^2) + D        (al - gl = D)
    //   =>  kt = Ta bNm1 D / Y,   ks = Ra + Ta t r bm1 / Y.
    // Both terms of Y are positive (al - r^2 = r (a - r) > 0), so nothing cancels -- the X form differences al^2 against r^2, which
    // loses digits as D -> 0 -- and three FP64 instructions per band and item go (12 instead of 15 after bNm1).
    const T Y = fma(bm1, al - r2, D);
    const T TaI = Ta * rcp_b(Y);                   // benign: scales kt and ks - Ra
    tran = TaI * (bNm1 * D);
    refl = fma(TaI * bm1, t * r, Ra);
#else
    const T dal = fma(al, al, -r2);                // al^2 - r^2 = r^2 (a^2 - 1)
    const T X = fma(bm1, al * (al - r2), dal);
    const T TaI = Ta * rcp_b(X);                   // benign: scales kt and ks - Ra
    tran = TaI * (bNm1 * dal);
    refl = fma(TaI * bm1, (t * al) * r, Ra);
#endif
    return hi_word(qmt) < 0x3EE00000;              // 1 - r - t < 2^-17 (r + t >= 1 included: negative high word), tested on the integer pipe
}
// Rare path of the double kernels: no or almost no absorption (1 - r - t < 2^-17).  IEEE divisions and libm throughout.
//   r + t >= 1: the zero-absorption case, exactly as models.py:1057-1065;
//   otherwise:  the Stokes formulas (:1043-1065) in the cancellation-free spelling of the fast path --
//     b - 1 = ((1 - r - t)(1 + r - t) + D)/(2t),  b^(2(N-1)) - 1 = expm1(2 (N-1) log1p(b - 1)),  Y = (b^(2(N-1)) - 1)(a r - r^2) + D --
//     which stays accurate to ~1e-15 down to D ~ 1e-8, where the reference itself carries 1e-16/D of rounding noise in a^2 - 1.
// (results by value -- .x = reflectance, .y = transmittance: reference parameters of a real call would force the caller's arrays into local memory)
template <typename T>
__device__ __noinline__ double2 leaf_layers_slow(T talf, T ralf, T t12, T r12, T t21, T r21, T tau, T nm1) {
    T refl, tran;
    const T one = T(1);
    const T x = r21 * tau;
    const T den1 = one - x * x;
    const T Ta = talf * tau * t21 / den1;
    const T Ra = ralf + x * Ta;
    const T t = t12 * tau * t21 / den1;
    const T r = r12 + x * t;
    if (r + t >= one) {                                               // :1057-1059
        const T Tsub = t / (t + (one - t) * nm1);
        const T Rsub = one - Tsub;
        const T den = one - Rsub * r;
        tran = Ta * Tsub / den;
        refl = Ra + Ta * Rsub * t / den;
        return make_double2(refl, tran);
    }
    const T q = one - r, qmt = q - t;
    const T D = sqrt(((one + r + t) * (one + r - t)) * ((q + t) * qmt));      // :1043
    const T al = T(0.5) * ((one + r * r - t * t) + D);                // a r   (:1046)
    const T lb = log1p((qmt * (one + r - t) + D) / (t + t));          // ln b  (:1047)
    const T bm1 = expm1((nm1 + nm1) * lb);                            // b^(2(N-1)) - 1
    const T bNm1 = exp(nm1 * lb);                                     // :1049
    const T Y = bm1 * (al - r * r) + D;
    tran = Ta * bNm1 * D / Y;                                         // :1062-1065 (see leaf_layers_fast)
    refl = Ra + Ta * t * r * bm1 / Y;
    return make_double2(refl, tran);
}

// ---- float (fp32 mode) variants of the plate + Stokes stage ----
// In float the reference's formulas lose everything near zero absorption: 1 - r - t is a difference of O(1) terms and
// a - 1, b - 1, b^(2(N-1)) - 1 are differences again.  This version is cancellation-free:
//   1 - r - t = t12 (1 - tau)/(1 - r21 tau)              (uses r21 + t21 = 1; 1 - tau = k g(k) from the table, relative 1e-7)
//   a r - r   = ((1 - r - t)(1 - r + t) + D)/2 ,   b t - t = ((1 - r - t)(1 + r - t) + D)/2
//   X = (a r)^2 b^(2(N-1)) - r^2 - a r r^2 (b^(2(N-1)) - 1) = da (al + r) + al (b^(2(N-1)) - 1)(al - r^2)   (all terms >= 0)
// Measured against the float64 oracle: <= 1e-6 for k >= 1e-4; below the table range (k < 2^-14, incl. k <= 0) the lane
// is flagged and re-evaluated in float64 by leaf_exact_f32().
__device__ __forceinline__ void leaf_layers_f32(const BandConst<float>& c, float omt, float nm1, float& refl, float& tran) {
    const float tau = 1.0f - omt;
    const float x = c.r21 * tau;
    const float inv = rcp(fmaf(-x, x, 1.0f));
    const float tt = tau * c.t21 * inv;
    const float Ta = c.talf * tt;
    const float t = c.t12 * tt;
    const float Ra = fmaf(x, Ta, c.ralf);
    const float r = fmaf(x, t, c.r12);
    const float p = 1.0f + r, q = 1.0f - r;
    const float qmt = (c.t12 * omt) * (fmaf(x, inv, inv));        // t12 (1 - tau) (1 + x)/(1 - x^2)
    const float P12 = (p + t) * (p - t);
    const float P34 = (q + t) * qmt;
    const float D = sqrt_pos(P12 * P34);
    const float da = 0.5f * (P34 + D);
    const float al = r + da;
    const float db = 0.5f * fmaf(qmt, p - t, D);
    const float b = fmaf(db, rcp(t), 1.0f);
    const float bNm1 = exp2_fast(fminf(log2_fast(b) * nm1, 60.0f));  // opaque plates: b^(N-1) saturates (its square must stay finite)
    const float bN2 = bNm1 * bNm1;
    const float bm1 = bN2 - 1.0f;
    const float dal = da * (al + r);                              // (a r)^2 - r^2
    const float X = fmaf(al * bm1, fmaf(-r, r, al), dal);
    const float TaI = Ta * rcp(X);
    tran = TaI * (bNm1 * dal);
    refl = fmaf(TaI * bm1, (t * al) * r, Ra);
}

// the same for the two bands of a thread at once, in packed FP32 (see math32.cuh)
__device__ __forceinline__ void leaf_layers_f32x2(const BandConst<f2>& c, f2 omt, float nm1, f2& refl, f2& tran) {
    const f2 tau = sub(F2(1.0f), omt);
    const f2 x = mul(c.r21, tau);
    const f2 inv = rcp(fma(neg(x), x, 1.0f));
    const f2 tt = mul(mul(tau, c.t21), inv);
    const f2 Ta = mul(c.talf, tt);
    const f2 t = mul(c.t12, tt);
    const f2 Ra = fma(x, Ta, c.ralf);
    const f2 r = fma(x, t, c.r12);
    const f2 p = add(r, 1.0f), q = sub(F2(1.0f), r);
    const f2 qmt = mul(mul(c.t12, omt), fma(x, inv, inv));        // t12 (1 - tau) (1 + x)/(1 - x^2)
    const f2 pmt = sub(p, t);
    const f2 P12 = mul(add(p, t), pmt);
    const f2 P34 = mul(add(q, t), qmt);
    const f2 D = sqrt_pos(mul(P12, P34));
    const f2 da = mul(add(P34, D), 0.5f);
    const f2 al = add(r, da);
    const f2 db = mul(fma(qmt, pmt, D), 0.5f);
    const f2 b = fma(db, rcp(t), 1.0f);
    const f2 lb = mul(log2_fast(b), nm1);
    const f2 bNm1 = exp2_fast(F2(fminf(lb.x, 60.0f), fminf(lb.y, 60.0f)));   // opaque plates: b^(N-1) saturates
    const f2 bN2 = mul(bNm1, bNm1);
    const f2 bm1 = add(bN2, -1.0f);
#if PB_F32_ALG
    // the Y form of leaf_layers_fast (round 2n): Y = bm1 (a r - r^2) + D, all terms >= 0; four packed instructions less
    const f2 Y = fma(bm1, fma(neg(r), r, al), D);
    const f2 TaI = mul(Ta, rcp(Y));
    tran = mul(TaI, mul(bNm1, D));
    refl = fma(mul(TaI, bm1), mul(t, r), Ra);
#else
    const f2 dal = mul(da, add(al, r));                           // (a r)^2 - r^2
    const f2 X = fma(mul(al, bm1), fma(neg(r), r, al), dal);
    const f2 TaI = mul(Ta, rcp(X));
    tran = mul(TaI, mul(bNm1, dal));
    refl = fma(mul(TaI, bm1), mul(mul(t, al), r), Ra);
#endif
}

// rare lanes of the float path (k outside [2^-14, 32)): the whole plate + Stokes stage in float64 with IEEE divisions and
// libm, exactly as models.py:953-957, 1019-1025, 1043-1065 (zero-absorption branch included)
// (arguments and results by value -- .x = reflectance, .y = transmittance: a reference parameter of a real call forces the caller's
//  variable into local memory, which put two STL.64 + two LDL.64 per item on the hot path of the packed fp32 instances until round 2n)
__device__ __noinline__ float2 leaf_exact_f32(float talf_, float ralf_, float t12_, float r12_, float t21_, float r21_, float kall, float nm1f) {
    const double tau = tau_rare((double)kall);
    const double t21 = t21_, r21 = r21_, t12 = t12_, r12 = r12_, talf = talf_, ralf = ralf_, nm1 = nm1f;
    const double x = r21 * tau;
    const double den1 = 1.0 - x * x;
    const double Ta = talf * tau * t21 / den1;
    const double Ra = ralf + x * Ta;
    const double t = t12 * tau * t21 / den1;
    const double r = r12 + x * t;
    if (!(t > 1e-280)) {                                          // opaque plate (k > ~640): nothing is transmitted
        return make_float2((float)Ra, 0.0f);
    }
    const double qmt = t12 * (1.0 - tau) / (1.0 - x);
    double Rsub, Tsub;
    if (!(qmt > 0.0)) {                                           // r + t >= 1                    :1057-1059
        Tsub = t / (t + (1.0 - t) * nm1);
        Rsub = 1.0 - Tsub;
    } else {
        const double D = sqrt((1.0 + r + t) * (1.0 + r - t) * (1.0 - r + t) * qmt);
        const double da = 0.5 * ((1.0 - r + t) * qmt + D);       // a r - r
        const double db = 0.5 * ((1.0 + r - t) * qmt + D);       // b t - t
        const double a = 1.0 + da / r;
        const double lb = log1p(db / t);
        const double bNm1 = exp(nm1 * lb);
        const double bm1 = expm1(2.0 * nm1 * lb);                 // b^(2(N-1)) - 1
        const double am1 = da / r * (2.0 + da / r);               // a^2 - 1
        const double den = am1 + a * a * bm1;                     // a^2 b^(2(N-1)) - 1
        if (den > 0.0) {
            Rsub = a * bm1 / den;                                 // :1052-1054
            Tsub = bNm1 * am1 / den;
        } else {
            Tsub = t / (t + (1.0 - t) * nm1);
            Rsub = 1.0 - Tsub;
        }
    }
    const double den = 1.0 - Rsub * r;
    return make_float2((float)(Ra + Ta * Rsub * t / den), (float)(Ta * Tsub / den));      // :1062-1065
}

// One-plate quantities exactly as models.py:1019-1025 (IEEE divisions), for the attributes PROSPECT.r/t/Ra/Ta/denom that the
// reference leaves behind (:959).  Only the run-time-selected PROSPECT instance calls this, when one of those planes is wanted.
template <typename T>
__device__ __noinline__ void leaf_one_layer(const BandConst<T>& c, T kall, T (&out)[5]) {
    const double tau = tau_full((double)kall);
    const double r21 = (double)c.r21, t21 = (double)c.t21;
    const double denom = 1.0 - r21 * r21 * tau * tau;
    const double Ta = (double)c.talf * tau * t21 / denom;
    const double Ra = (double)c.ralf + r21 * tau * Ta;
    const double t = (double)c.t12 * tau * t21 / denom;
    const double r = (double)c.r12 + r21 * tau * t;
    out[0] = (T)r; out[1] = (T)t; out[2] = (T)Ra; out[3] = (T)Ta; out[4] = (T)denom;
}

// Stage 1 of the double leaf path: kall and the table value tau(kall) of all NU bands -- branch free; lanes outside the table range are
// flagged (bit u of `rare`) and repaired by leaf_tau_repair() before stage 2.  Split off so that the item loop can evaluate it for the
// NEXT item inside the canopy stage of the current one (software pipelining, PB_PIPELINE): it is the load- and integer-heavy part
// of the leaf (25 LDS and 50 integer / conversion instructions for 30 FP64 ones), and in the same basic block as ~270 FP64
// instructions it fills issue slots that the FP64 pipe leaves empty.
template <typename CF, bool CAN, typename T, typename MT>
__device__ __forceinline__ void leaf_tau(const T* __restrict__ kc, const T* __restrict__ s, int kz, const MT& mt, T (&tau)[CF::NU], unsigned& rare) {
    T kall[CF::NU];
    leaf_kall_all<CF, CAN>(kc, s, kz, kall);
    rare = 0;
#pragma unroll
    for (int u = 0; u < CF::NU; ++u) {
        bool r;
        tau[u] = tau_fast(kall[u], mt.tau, r);                        // :953-957
        rare |= r ? (1u << u) : 0u;
    }
}
template <typename CF, bool CAN, typename T>
__device__ __forceinline__ void leaf_tau_repair(const T* __restrict__ kc, const T* __restrict__ s, int kz, T (&tau)[CF::NU], unsigned rare) {
    if (__builtin_expect(rare != 0, 0)) {
        T kall[CF::NU];
        leaf_kall_all<CF, CAN>(kc, s, kz, kall);                               // recomputed: the rare path does not keep kall alive across the pipeline
#pragma unroll
        for (int u = 0; u < CF::NU; ++u)
            if ((rare >> u) & 1u) tau[u] = tau_rare(kall[u]);
    }
}
// Stage 2: plate + Stokes for all NU bands from tau (double)
template <typename CF, bool CSMEM, typename T, typename MT>
__device__ __forceinline__ void leaf_layers_all(const BandConst<T> (&c)[CF::NU], const T* __restrict__ kc, const T* __restrict__ s, const MT& mt,
                                                const T (&tau)[CF::NU], T (&refl)[CF::NU], T (&tran)[CF::NU]) {
    constexpr int NU = CF::NU;
    const T nm1 = s[S_NM1];
    {
        bool zero[NU];
        bool any_zero = false;
        const int lmax_hi = hi_word(s[S_LMAX]);
        T ta21[NU], ralf[NU], t1221[NU], r12[NU];
        if constexpr (CSMEM) {
            static_assert(KC_TA21 == 8 && KC_RALF == 9 && KC_T1221 == 10 && KC_R12 == 11, "plane pairs");
            load_pair<CF>(kc, KC_TA21, ta21, ralf);
            load_pair<CF>(kc, KC_T1221, t1221, r12);
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            if constexpr (CSMEM)
                zero[u] = leaf_layers_fast(kc[u + CF::koff(KC_R21)], ta21[u], t1221[u], ralf[u], r12[u],
                                           tau[u], nm1, lmax_hi, mt, refl[u], tran[u]);
            else
                zero[u] = leaf_layers_fast(c[u].r21, c[u].ta21, c[u].t1221, c[u].ralf, c[u].r12, tau[u], nm1, lmax_hi, mt, refl[u], tran[u]);
            any_zero |= zero[u];
        }
        if (__builtin_expect(any_zero, 0)) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (zero[u]) {
                    if constexpr (CSMEM) {       // talf = 1 - (1 - talf) etc.: within an ulp of the uploaded constants
                        const T ralf = kc[u + CF::koff(KC_RALF)], r12 = kc[u + CF::koff(KC_R12)], r21 = kc[u + CF::koff(KC_R21)];
                        const double2 rt = leaf_layers_slow(T(1) - ralf, ralf, T(1) - r12, r12, T(1) - r21, r21, tau[u], nm1);
                        refl[u] = rt.x; tran[u] = rt.y;
                    } else {
                        const double2 rt = leaf_layers_slow(c[u].talf, c[u].ralf, c[u].t12, c[u].r12, c[u].t21, c[u].r21, tau[u], nm1);
                        refl[u] = rt.x; tran[u] = rt.y;
                    }
                }
        }
    }
}

// all NU bands of a thread: the chains are independent and written back to back so that they interleave.
// CSMEM (double, fused PROSAIL instances): the interface constants are read from the thread's columns of SailSmem::kc instead of
// c[] -- 20 registers less across the item loop; the rare zero-absorption path rebuilds the plain constants from the complements.
template <typename CF, bool CSMEM, bool CAN, typename T, typename MT>
__device__ __forceinline__ void leaf_optics(const BandConst<T> (&c)[CF::NU], const T* __restrict__ kc, const T* __restrict__ s, int kz, const MT& mt,
                                            T (&refl)[CF::NU], T (&tran)[CF::NU]) {
    constexpr int NU = CF::NU;
    T kall[NU];
    bool rare[NU];
    bool any_rare = false;
    const T nm1 = s[S_NM1];
    leaf_kall_all<CF, CAN>(kc, s, kz, kall);
    if constexpr (std::is_same<T, float>::value) {
        float omt[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) omt[u] = kall[u] * tau_fast(kall[u], mt.tau, rare[u]);     // 1 - tau = k g(k)
#pragma unroll
        for (int u = 0; u < NU; ++u) { leaf_layers_f32(c[u], omt[u], nm1, refl[u], tran[u]); any_rare |= rare[u]; }
        if (__builtin_expect(any_rare, 0)) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (rare[u]) {
                    const float2 rt = leaf_exact_f32(c[u].talf, c[u].ralf, c[u].t12, c[u].r12, c[u].t21, c[u].r21, kall[u], nm1);
                    refl[u] = rt.x; tran[u] = rt.y;
                }
        }
    } else {
        T tau[NU];
        unsigned rmask = 0;
#pragma unroll
        for (int u = 0; u < NU; ++u) { tau[u] = tau_fast(kall[u], mt.tau, rare[u]); rmask |= rare[u] ? (1u << u) : 0u; }   // :953-957
        leaf_tau_repair<CF, CAN>(kc, s, kz, tau, rmask);
        leaf_layers_all<CF, CSMEM>(c, kc, s, mt, tau, refl, tran);
    }
}

// fp32 mode, two bands per thread: kall and the table look-ups per band (scalar), plate + Stokes stage packed
template <typename CF, bool CAN, typename MT>
__device__ __forceinline__ void leaf_optics_x2(const BandConst<f2>& c2, const float* __restrict__ kc, const float* __restrict__ s, int kz, const MT& mt,
                                               f2& refl, f2& tran) {
    static_assert(CF::NU == 2, "the packed path pairs the two bands of a thread");
    const float nm1 = s[S_NM1];
    float pa[2], pb[2];
    f2 kall;                                                                       // :950-951, both bands per instruction; kz: see leaf_kall_all
    if (kz == 2) {
        load_pair<CF>(kc, KC_KW, pa, pb);
        kall = fma(s[S_CM], F2(pb[0], pb[1]), mul(F2(pa[0], pa[1]), s[S_CWX]));
    } else {
        if (kz == 1) {
            load_pair<CF>(kc, KC_KAN, pa, pb);                                     // one 128-bit load per pair of planes
            kall = mul(F2(pb[0], pb[1]), s[S_CBRX]);
        } else {
            load_pair<CF>(kc, KC_KAB, pa, pb);
            kall = fma(s[S_CXC], F2(pb[0], pb[1]), mul(F2(pa[0], pa[1]), s[S_CAB]));
            load_pair<CF>(kc, KC_KAN, pa, pb);
            kall = fma(s[S_CBR], F2(pb[0], pb[1]), kall);
        }
        if constexpr (CAN) kall = fma(s[S_CAN], F2(pa[0], pa[1]), kall);
        load_pair<CF>(kc, KC_KW, pa, pb);
        kall = fma(s[S_CM], F2(pb[0], pb[1]), fma(s[S_CW], F2(pa[0], pa[1]), kall));
    }
    const float k0 = kall.x, k1 = kall.y;
    bool rare0, rare1;
    const float g0 = tau_fast(k0, mt.tau, rare0), g1 = tau_fast(k1, mt.tau, rare1);
    leaf_layers_f32x2(c2, mul(kall, F2(g0, g1)), nm1, refl, tran);                 // 1 - tau = k g(k)
    if (__builtin_expect(rare0 | rare1, 0)) {
        if (rare0) {
            const float2 rt = leaf_exact_f32(c2.talf.x, c2.ralf.x, c2.t12.x, c2.r12.x, c2.t21.x, c2.r21.x, k0, nm1);
            refl.x = rt.x; tran.x = rt.y;
        }
        if (rare1) {
            const float2 rt = leaf_exact_f32(c2.talf.y, c2.ralf.y, c2.t12.y, c2.r12.y, c2.t21.y, c2.r21.y, k1, nm1);
            refl.y = rt.x; tran.y = rt.y;
        }
    }
}

// geometry-independent part of 4SAIL for one band (models.py:590-601, 637-642, 654-655, 714-719)
template <typename T>
struct Diffuse {
    T m, e1, rinf, re, invden, inv_omr2, tdd, rdd, rs_invdn, rsrdd, rinf_invden;
    T invdn;       // 1/(1 - rs rdd)
    T invden_h;    // PB_HALVES (AB form): invden/2, and inv_omr2 then holds 1/(4 (1 - rinf^2)) -- see sail_diffuse
    T a1, a2;      // rho + tau rinf, tau + rho rinf: the leaf-dependent factors of (sf + sb rinf), (sf rinf + sb), ... (:649-652)
                   // fused fp64 instances (AB form): a1 = Ap = (a1 + a2)/2 = (rho + tau)(1 + rinf)/2,  a2 = Am = bf (rho - tau)(1 - rinf)/2,
                   // so that  sf + sb rinf = k Ap - Am,  sf rinf + sb = k Ap + Am  (and the same with K): one FMA each per geometry
};

// FROM_LEAF: rho, tau come from the PROSPECT stage of the same kernel.  Then 0 < rho, tau and rho + tau < 1 hold by
// construction (rho >= the interface reflectance, absorption > 0 or the zero-absorption branch), so sigb, sigf are never
// zero (the 1e-36 floors of models.py:593-598 cannot fire) and att^2 - sigb^2 = (1 - rho - tau)(...) stays positive unless
// rho + tau rounds to 1, which sqrt_pos still guards.  With spectra supplied by the caller (SAIL mode) the floors are kept.
template <bool FROM_LEAF, typename T, typename MT>
__device__ __forceinline__ void sail_diffuse(T rho, T tau, T rs, const T* __restrict__ s, const MT& mt, Diffuse<T>& d) {
    const T one = T(1), tiny = T(1e-36);
    if constexpr (FROM_LEAF && !is_f32<T>()) {
        // With s = rho + tau, dd = rho - tau and ddb - ddf = bf (:587-588):  sigb = (s + bf dd)/2,  sigf = (s - bf dd)/2,
        //   att - sigb = 1 - s,   att + sigb = 1 + bf dd   =>   m^2 = att^2 - sigb^2 = (1 - s)(1 + bf dd)   (:590-601, no cancellation)
        //   rinf = sigb/(att + m) = (s + bf dd) / (2 - (s - bf dd) + 2 m)                                    (:639)
        const T s_ = rho + tau, dd = rho - tau;
        const T bd = s[S_BF] * dd;
        const T oms = one - s_;
#if PB_RINF_P
        // Round 2n: with P = att + sigb = 1 + bd and Q = att - sigb = 1 - s:  m = sqrt(P Q)  and
        //   rinf = (att - m)/sigb = (P - m)/(P + m)        [(att - m)(P + m) - sigb (P - m) = P Q - m^2 = 0]
        // two FP64 instructions less than sigb/(att + m) spelled out in s, bd.  P - m cancels only where rinf itself is small
        // (dark leaves): the second-order square root (3e-12 relative in m) then leaves <= 1.5e-12 ABSOLUTE in rinf.
        const T P = one + bd;
        // rho + tau >= 1 (a leaf without absorption, to rounding): the reference's m is 0 or NaN there.  The argument is floored at the
        // smallest normal number by ONE integer max on its high word (m ~ 1e-154: every later formula sees m = 0) instead of a
        // compare and two selects on the result.
        d.m = sqrt_b_raw(floor_hi(oms * P, 0x00100000));
        d.rinf = (P - d.m) * rcp_b(P + d.m);
#else
        d.m = sqrt_b_pos(fma(oms, bd, oms));
        d.rinf = (s_ + bd) * rcp_b(fma(T(2), d.m, T(2) - (s_ - bd)));
#endif
#if PB_HALVES
        // 2 Ap and 2 Am: one FMA each.  Everything downstream is homogeneous in them -- tsd, rsd, tdo, rdo of degree one, T1 + T2 - T3
        // of degree two -- so the factors 1/2 and 1/4 are taken out of invden (invden_h) and inv_omr2 below: two exponent
        // decrements on the integer pipe instead of two FP64 instructions.
        d.a1 = fma(s_, d.rinf, s_);
        d.a2 = fma(-bd, d.rinf, bd);
#else
        d.a1 = s_ * fma(T(0.5), d.rinf, T(0.5));                      // Ap
        d.a2 = bd * fma(T(-0.5), d.rinf, T(0.5));                     // Am
#endif
    } else {
    T sigb = fma(s[S_DDB], rho, s[S_DDF] * tau);                      // :590
    T sigf = fma(s[S_DDF], rho, s[S_DDB] * tau);                      // :591
    if (!FROM_LEAF) {
        sigb = is_zero(sigb) ? tiny : sigb;                           // :594-595 (tests on the integer pipe)
        sigf = is_zero(sigf) ? tiny : sigf;
    }
    const T att = one - sigf;                                         // :600
    if constexpr (std::is_same<T, float>::value) {
        // float: m^2 = (att - sigb)(att + sigb) and rinf = sigb/(att + m) avoid the two cancellations of :601, :639
        d.m = sqrt_pos((att - sigb) * (att + sigb));
        d.rinf = sigb * rcp(att + d.m);
    } else {
        d.m = sqrt_b_pos(fma(att, att, -(sigb * sigb)));              // :601
#if PB_BENIGN_Q
        // :639 as sigb / (att + m) (= (att - m)/sigb because m^2 = att^2 - sigb^2): a relative error of m then reaches rinf
        // divided by (att + m) >= 1 instead of multiplied by m/sigb -- what makes the second-order sqrt benign here
        d.rinf = sigb * rcp_b(att + d.m);
#else
        d.rinf = (att - d.m) * rcp_k(sigb);                           // :639
#endif
    }
    d.a1 = fma(tau, d.rinf, rho);
    d.a2 = fma(rho, d.rinf, tau);
    }
#if PB_EXPS
    if constexpr (!is_f32<T>()) d.e1 = exp_neg_scaled(d.m, s[S_NEGLAI], mt.exp2t);   // the double record carries -lai 2048 log2(e)      :637
    else
#endif
    d.e1 = exp_neg(d.m * s[S_NEGLAI], mt);                            // :637
    d.re = d.rinf * d.e1;
    const T omr2 = fma(-d.rinf, d.rinf, one);                         // 1 - rinf^2
    const T den = fma(-d.re, d.re, one);                              // 1 - rinf^2 e^2 = 1 - (rinf e)^2          :642
    if constexpr (is_f32<T>() || !(PB_QUADRATIC || PB_UNMERGE)) {
        // 1/(1 - rinf^2 e^2) (:642) and 1/(1 - rinf^2) (:678) from ONE reciprocal of their product: one MUFU less (an SFU
        // instruction next to a busy FP64 pipe costs about as much as the two multiplications that replace it, and the
        // unmerged build measured 5% slower: profiles/r2_variants.txt).  Both results only scale outputs: benign.
        const T r = rcp_b(den * omr2);
        d.invden = r * omr2;
        d.inv_omr2 = r * den;
    } else {                                                          // two 2-instruction reciprocals (experiment)
        d.invden = rcp_b(den);
        d.inv_omr2 = rcp_b(omr2);
    }
#if PB_HALVES
    if constexpr (FROM_LEAF && !is_f32<T>()) {                        // both are >= 1: the exponent decrements are exact
        d.invden_h = half_of(d.invden);
        d.inv_omr2 = quarter_of(d.inv_omr2);
    }
#endif
    d.tdd = omr2 * d.e1 * d.invden;                                   // :654
    d.rinf_invden = d.rinf * d.invden;                                // shared by rdd and by T3 of the directional stage
    d.rdd = fma(-d.re, d.e1, d.rinf) * d.invden;                      // rinf (1 - e^2)/den = (rinf - (rinf e) e)/den   :655
    d.rsrdd = rs * d.rdd;
    T dn = one - d.rsrdd;                                             // :714-719
    if constexpr (!is_f32<T>()) dn = floor_hi(dn, 0x38754484);        // dn < 1e-36, negatives too -> 1e-36 (to 2^-20 relative): one integer max on the high word
    else dn = below_tiny(dn) ? tiny : dn;                             // dn < 1e-36 (integer-pipe test; negatives too)
    d.invdn = rcp_b(dn);
    d.rs_invdn = rs * d.invdn;
}

template <typename T>
struct Canopy {
    T rsot, rddt, rsdt, rdot;        // BRF, BHR, DHR, HDR with soil
    T rso, rsd, tsd, rdo, tdo;       // canopy-only terms
};

// common tail of the directional part once J1(k,m), J1(K,m), 1/(k+m), 1/(K+m) are known
// (models.py:603-607, 649-678, 706-726)
template <bool AB, typename T>
__device__ __forceinline__ void sail_couple(T rho, T tau, T rs, const Diffuse<T>& d, const T* __restrict__ g, T J1ks, T J1ko, T inv_kp,
                                            T inv_Kp, unsigned need, Canopy<T>& o) {
    const T one = T(1);
    const T tss = g[G_TSS], too = g[G_TOO];
    // sb, sf, vb, vf (:603-606) only ever appear as sf + sb rinf, sf rinf + sb, vf + vb rinf, vf rinf + vb (:649-652).  With
    // a1 = rho + tau rinf and a2 = tau + rho rinf (geometry independent, computed once per parameter set):
    //   sf + sb rinf = sdf a1 + sdb a2,   sf rinf + sb = sdf a2 + sdb a1,   and the same with (dof, dob) for the view direction
    T u1, u2, v1, v2;
    if constexpr (AB) {           // d.a1 = Ap, d.a2 = Am (see sail_diffuse): one FMA each
        u1 = fma(g[G_K], d.a1, -d.a2); u2 = fma(g[G_K], d.a1, d.a2);
        v1 = fma(g[G_KO], d.a1, -d.a2); v2 = fma(g[G_KO], d.a1, d.a2);
    } else {
        u1 = fma(g[G_SDF], d.a1, g[G_SDB] * d.a2); u2 = fma(g[G_SDF], d.a2, g[G_SDB] * d.a1);
        v1 = fma(g[G_DOF], d.a1, g[G_DOB] * d.a2); v2 = fma(g[G_DOF], d.a2, g[G_DOB] * d.a1);
    }
    // J2(k,m) = (1 - e1 tss)/(k+m), J2(K,m) = (1 - e1 too)/(K+m) (:646-647) only appear multiplied by u2 resp. v2, and so do the
    // 1/(k+m), 1/(K+m) of g2, g1 (:668-669): ck = u2/(k+m) and cK = v2/(K+m) are formed once
    const T ck = u2 * inv_kp, cK = v2 * inv_Kp;
    const T Pss = u1 * J1ks, Qss = ck * fma(-d.e1, tss, one), Pv = v1 * J1ko, Qv = cK * fma(-d.e1, too, one);
    // tsd, rsd, tdo, rdo (:656-659) before their common factor 1/(1 - rinf^2 e^2): T3 takes it together with rinf
    const T tsd_u = fma(-d.re, Qss, Pss), tdo_u = fma(-d.re, Qv, Pv), rdo_u = fma(-d.re, Pv, Qv);
    // (PB_HALVES, AB form: u1 .. v2 are twice the reference's, the four terms take invden/2; T1 + T2 - T3 is four times too large
    //  and inv_omr2 carries the 1/4)
    T invden1;
    if constexpr (AB && PB_HALVES) invden1 = d.invden_h; else invden1 = d.invden;
    o.tsd = tsd_u * invden1;
    o.rsd = fma(-d.re, Pss, Qss) * invden1;
    o.tdo = tdo_u * invden1;
    o.rdo = rdo_u * invden1;
    // T1 + T2 - T3 as nested FMAs (:668-678):  T1 = v2 g1 u1 with g1 = (z - J1(k,m) too)/(K+m),  T2 = v1 g2 u2,
    // T3 = (rdo Qss + tdo Pss) rinf
    const T T3 = fma(rdo_u, Qss, tdo_u * Pss) * d.rinf_invden;
    const T Ts = fma(cK * u1, fma(-J1ks, too, g[G_Z]), fma(ck * v1, fma(-J1ko, tss, g[G_Z]), -T3));
    // rsos + rsod apart from Ts/(1 - rinf^2): w lai sumint with w = Fs rho + Ft tau (:607, :706); fused fp64 instances: the prologue
    // folds lai sumint into Fs, Ft (G_FSL, G_FTL)
    if constexpr (!is_f32<T>() && PB_SOILFOLD) {
        // Soil coupling (:721-726), round 2n.  With A1 = tss + tsd and A2 = too + tdo (total down- and upward transmittances):
        //   (tss + tsd) tdo + (tsd + tss rs rdd) too = A1 A2 - tss too (1 - rs rdd) = A1 A2 - tss too dn
        //   =>  rsot = rsod + rsos + tsstoo rs + rsodt = rsod + rsos + [A1 A2 / dn + (tsstoo - tss too)] rs
        // tsstoo - tss too comes from the prologue (G_DTT); A1, A2 are ONE FMA each on the unscaled numerators.  8 FP64
        // instructions for the BRF instead of 12, all errors absolute at the 1e-16 level (the subtraction happens in the record).
        const T A1 = fma(tsd_u, invden1, tss), A2 = fma(tdo_u, invden1, too);
        const T cpl = fma(A1 * A2, d.invdn, g[G_DTT]);
        if (need & (1u << O_RSO)) {
            if constexpr (AB && PB_FSFOLD) o.rso = fma(g[G_FSL], rho, fma(g[G_FTL], tau, Ts * d.inv_omr2));     // rsos + rsod    :706, :710
            else o.rso = fma(fma(g[G_FS], rho, g[G_FT] * tau), g[G_LAISUM], Ts * d.inv_omr2);
            o.rsot = fma(cpl, rs, o.rso);
        } else {
            T ss;
            if constexpr (AB && PB_FSFOLD) ss = fma(g[G_FSL], rho, g[G_FTL] * tau);
            else ss = fma(g[G_FS], rho, g[G_FT] * tau) * g[G_LAISUM];
            o.rsot = fma(cpl, rs, fma(Ts, d.inv_omr2, ss));
        }
        if (need & (1u << O_RDDT)) o.rddt = fma(d.tdd * d.tdd, d.rs_invdn, d.rdd);
        if (need & (1u << O_RSDT)) o.rsdt = fma(A1 * d.tdd, d.rs_invdn, o.rsd);
        if (need & (1u << O_RDOT)) o.rdot = fma(d.tdd * A2, d.rs_invdn, o.rdo);
    } else {
        // soil coupling (:721-726): rsodt = ((tss + tsd) tdo + (tsd + tss rs rdd) too) rs/dn
        const T sc = fma(tss + o.tsd, o.tdo, fma(tss, d.rsrdd, o.tsd) * too);
        const T w = fma(g[G_FS], rho, g[G_FT] * tau);                 // :607
        if (need & (1u << O_RSO)) {
            o.rso = fma(w, g[G_LAISUM], Ts * d.inv_omr2);             // rsos + rsod                              :706, :710
            o.rsot = fma(sc, d.rs_invdn, fma(g[G_TSSTOO], rs, o.rso));
        } else {                                                      // rsot = rsod + rsos + tsstoo rs + rsodt, one FMA per term
            o.rsot = fma(sc, d.rs_invdn, fma(Ts, d.inv_omr2, fma(w, g[G_LAISUM], g[G_TSSTOO] * rs)));
        }
        if (need & (1u << O_RDDT)) o.rddt = fma(d.tdd * d.tdd, d.rs_invdn, d.rdd);
        if (need & (1u << O_RSDT)) o.rsdt = fma((o.tsd + tss) * d.tdd, d.rs_invdn, o.rsd);
        if (need & (1u << O_RDOT)) o.rdot = fma(d.tdd * (o.tdo + too), d.rs_invdn, o.rdo);
    }
}

// Branch-free geometry-dependent part of 4SAIL (models.py:603-607, 644-678, 706-726, 765-789).
// double: returns true when one of the J1 terms needs its series form (|k - m| lai <= 1e-3, models.py:777-779); the
// caller then re-evaluates that band with sail_directional().
// float: J1(k,m) = (e^{-m t} - e^{-k t})/(k - m) = e^{-k t} t expm1(d)/d with d = (k - m) t has no cancellation for any d,
// so there is no series branch and the function never returns true.
template <bool AB, typename T>
__device__ __forceinline__ bool sail_directional_fast(T rho, T tau, T rs, const Diffuse<T>& d, const T* __restrict__ s,
                                                      const T* __restrict__ g, int thr_hi, unsigned need, Canopy<T>& o) {
    const T k = g[G_K], K = g[G_KO], tss = g[G_TSS], too = g[G_TOO];
    const T km = k - d.m, kp = k + d.m, Km = K - d.m, Kp = K + d.m;
    if constexpr (std::is_same<T, float>::value) {
        const T lai = s[S_LAI];
        // direct difference where |d| >= 0.25 (at most two bits lost), series below: see sail_directional_f32x2
        const T xk = km * lai, xK = Km * lai;
        const T J1ks = fabsf(xk) < 0.25f ? (tss * lai) * expm1_over_x(xk) : (d.e1 - tss) * rcp(km);
        const T J1ko = fabsf(xK) < 0.25f ? (too * lai) * expm1_over_x(xK) : (d.e1 - too) * rcp(Km);
        sail_couple<AB>(rho, tau, rs, d, g, J1ks, J1ko, rcp(kp), rcp(Kp), need, o);
        return false;
    } else {
        // 1/(k-m) and 1/(k+m) from ONE reciprocal of (k-m)(k+m)                             :644-647, 765-789
        const bool rare = !(abs_gt_hi(km, thr_hi) && abs_gt_hi(Km, thr_hi));
#if PB_QUADRATIC || PB_UNMERGE
        // four 2-instruction reciprocals (8 FP64-pipe instructions + 4 MUFU) instead of the merged form below (11 + 1 MUFU):
        // measured SLOWER (round 2b) -- the three extra SFU instructions cost more than the three FP64 instructions saved
        const T J1ks = (d.e1 - tss) * rcp_b(km);
        const T J1ko = (d.e1 - too) * rcp_b(Km);
        sail_couple<AB>(rho, tau, rs, d, g, J1ks, J1ko, rcp_b(kp), rcp_b(Kp), need, o);
#else
        // ... and both of those from one reciprocal of (k-m)(k+m)(K-m)(K+m)
        // all four reciprocals only scale their numerators (the difference e1 - tss is not touched): benign
        const T pk = km * kp, pK = Km * Kp;
        const T rr = rcp_b(pk * pK);
        const T ik = rr * pK, iK = rr * pk;
        const T J1ks = (d.e1 - tss) * (ik * kp);
        const T J1ko = (d.e1 - too) * (iK * Kp);
        sail_couple<AB>(rho, tau, rs, d, g, J1ks, J1ko, ik * km, iK * Km, need, o);
#endif
        return rare;
    }
}

// geometry-dependent part of 4SAIL for one band with J1's series branch (models.py:765-785): the repair path (double)
template <bool AB, typename T>
__device__ __forceinline__ void sail_directional(T rho, T tau, T rs, const Diffuse<T>& d, const T* __restrict__ s,
                                                 const T* __restrict__ g, int thr_hi, unsigned need, Canopy<T>& o) {
    const T k = g[G_K], K = g[G_KO], tss = g[G_TSS], too = g[G_TOO], lai = s[S_LAI];
    const T half = T(0.5), one = T(1);
    // J1(k,m), J2(k,m), J1(K,m), J2(K,m): 1/(k-m) and 1/(k+m) from ONE reciprocal of (k-m)(k+m)   :644-647, 765-789
    T J1ks, J1ko, inv_kp, inv_Kp;
    {
        const T km = k - d.m, kp = k + d.m;
        if (__builtin_expect(abs_gt_hi(km, thr_hi), 1)) {
            const T i = rcp(km * kp);
            inv_kp = i * km;
            J1ks = (d.e1 - tss) * (i * kp);
        } else {
            const T del = km * lai;
            inv_kp = rcp(kp);
            J1ks = half * lai * (tss + d.e1) * (one - del * del / T(12));
        }
        const T Km = K - d.m, Kp = K + d.m;
        if (__builtin_expect(abs_gt_hi(Km, thr_hi), 1)) {
            const T i = rcp(Km * Kp);
            inv_Kp = i * Km;
            J1ko = (d.e1 - too) * (i * Kp);
        } else {
            const T del = Km * lai;
            inv_Kp = rcp(Kp);
            J1ko = half * lai * (too + d.e1) * (one - del * del / T(12));
        }
    }
    sail_couple<AB>(rho, tau, rs, d, g, J1ks, J1ko, inv_kp, inv_Kp, need, o);
}

// ---------------------------------------------------------------------------------------------
// 4SAIL for the two bands of a thread in packed FP32 (fp32 mode).  Same formulas as sail_diffuse<float>,
// sail_directional_fast<float> and sail_couple<float>; per-item scalars enter as broadcast operands.
// ---------------------------------------------------------------------------------------------
template <bool FROM_LEAF>
__device__ __forceinline__ void sail_diffuse_f32x2(f2 rho, f2 tau, f2 rs, const float* __restrict__ s, Diffuse<f2>& d) {
    if constexpr (FROM_LEAF && PB_F32_ALG) {
        // the P / Q form of sail_diffuse (round 2n): P = att + sigb = 1 + bf (rho - tau), Q = att - sigb = 1 - rho - tau,
        // m = sqrt(P Q), rinf = (P - m)/(P + m); a1 = 2 Ap = s (1 + rinf), a2 = Am = (bf/2)(rho - tau)(1 - rinf): the directional
        // stage multiplies a1 with k/2, K/2 (scalars, once per item)
        const f2 s_ = add(rho, tau), dd = sub(rho, tau);
        const f2 oms = sub(F2(1.0f), s_);
        const f2 P = fma(s[S_BF], dd, F2(1.0f));
        const f2 bdh = mul(dd, 0.5f * s[S_BF]);
        d.m = sqrt_pos(mul(oms, P));
        d.rinf = mul(sub(P, d.m), rcp(add(P, d.m)));
        d.a1 = fma(s_, d.rinf, s_);
        d.a2 = fma(neg(bdh), d.rinf, bdh);
    } else {
        f2 sigb = fma(s[S_DDB], rho, mul(tau, s[S_DDF]));                 // :590
        f2 sigf = fma(s[S_DDF], rho, mul(tau, s[S_DDB]));                 // :591
        if (!FROM_LEAF) {
            sigb = F2(is_zero(sigb.x) ? 1e-36f : sigb.x, is_zero(sigb.y) ? 1e-36f : sigb.y);      // :594-595
            sigf = F2(is_zero(sigf.x) ? 1e-36f : sigf.x, is_zero(sigf.y) ? 1e-36f : sigf.y);
        }
        const f2 att = sub(F2(1.0f), sigf);                               // :600
        d.m = sqrt_pos(mul(sub(att, sigb), add(att, sigb)));              // :601 without the cancellation
        d.rinf = mul(sigb, rcp(add(att, d.m)));                           // :639 as sigb / (att + m)
        d.a1 = fma(tau, d.rinf, rho);
        d.a2 = fma(rho, d.rinf, tau);
    }
    d.e1 = exp2_fast(mul(d.m, s[S_NEGLAI] * 1.4426950408889634f));    // :637
    d.re = mul(d.rinf, d.e1);
    const f2 omr2 = fma(neg(d.rinf), d.rinf, 1.0f);                   // 1 - rinf^2
    const f2 den = fma(neg(d.re), d.re, 1.0f);                        // 1 - (rinf e)^2                        :642
    const f2 r = rcp(mul(den, omr2));
    d.invden = mul(r, omr2);
    d.inv_omr2 = mul(r, den);
    d.tdd = mul(mul(omr2, d.e1), d.invden);                           // :654
    d.rinf_invden = mul(d.rinf, d.invden);
    d.rdd = mul(fma(neg(d.re), d.e1, d.rinf), d.invden);              // rinf (1 - e^2)/den                    :655
    d.rsrdd = mul(rs, d.rdd);
    f2 dn = sub(F2(1.0f), d.rsrdd);                                   // :714-719
    dn = F2(below_tiny(dn.x) ? 1e-36f : dn.x, below_tiny(dn.y) ? 1e-36f : dn.y);
    d.invdn = rcp(dn);
    d.rs_invdn = mul(rs, d.invdn);
}

template <bool AB>
__device__ __forceinline__ void sail_directional_f32x2(f2 rho, f2 tau, f2 rs, const Diffuse<f2>& d, const float* __restrict__ s,
                                                       const float* __restrict__ g, unsigned need, Canopy<f2>& o) {
    const float k = g[G_K], K = g[G_KO], tss = g[G_TSS], too = g[G_TOO], lai = s[S_LAI];
    const f2 km = sub(F2(k), d.m), kp = add(d.m, k), Km = sub(F2(K), d.m), Kp = add(d.m, K);
    // J1(k,m) = (e^{-m t} - e^{-k t})/(k - m)  (:765-785).  Where |d| = |k - m| t >= 0.25 the difference loses at most two bits
    // and is taken directly; below, J1 = e^{-k t} t expm1(d)/d with a series: no cancellation anywhere, no repair path, and no
    // intermediate can overflow or turn 0 * inf into NaN when tss or e1 underflow (lai of hundreds, grazing sun).
    const f2 xk = mul(km, lai), xK = mul(Km, lai);
    const f2 sk = mul(expm1_over_x_small(xk), tss * lai), sK = mul(expm1_over_x_small(xK), too * lai);
    const f2 dk = mul(add(d.e1, -tss), rcp(km)), dK = mul(add(d.e1, -too), rcp(Km));
    const f2 J1ks = F2(fabsf(xk.x) < 0.25f ? sk.x : dk.x, fabsf(xk.y) < 0.25f ? sk.y : dk.y);
    const f2 J1ko = F2(fabsf(xK.x) < 0.25f ? sK.x : dK.x, fabsf(xK.y) < 0.25f ? sK.y : dK.y);
    const f2 rkK = rcp(mul(kp, Kp));                                  // 1/(k+m) and 1/(K+m) from one reciprocal (both > 0)
    const f2 inv_kp = mul(rkK, Kp), inv_Kp = mul(rkK, kp);
    f2 u1, u2, v1, v2;
    if constexpr (AB) {           // d.a1 = 2 Ap, d.a2 = Am (sail_diffuse_f32x2): sf + sb rinf = k Ap - Am, sf rinf + sb = k Ap + Am, same with K
        const float kh = 0.5f * k, Kh = 0.5f * K;
        u1 = fma(kh, d.a1, neg(d.a2)); u2 = fma(kh, d.a1, d.a2);
        v1 = fma(Kh, d.a1, neg(d.a2)); v2 = fma(Kh, d.a1, d.a2);
    } else {
        u1 = fma(g[G_SDF], d.a1, mul(d.a2, g[G_SDB])); u2 = fma(g[G_SDF], d.a2, mul(d.a1, g[G_SDB]));   // :603-606, 649-652 via a1, a2
        v1 = fma(g[G_DOF], d.a1, mul(d.a2, g[G_DOB])); v2 = fma(g[G_DOF], d.a2, mul(d.a1, g[G_DOB]));
    }
    const f2 ck = mul(u2, inv_kp), cK = mul(v2, inv_Kp);             // u2/(k+m), v2/(K+m): shared by Q (J2, :787-789) and T2, T1
    const f2 Pss = mul(u1, J1ks), Qss = mul(ck, fma(neg(d.e1), F2(tss), 1.0f)), Pv = mul(v1, J1ko), Qv = mul(cK, fma(neg(d.e1), F2(too), 1.0f));
    const f2 nre = neg(d.re);
    const f2 tsd_u = fma(nre, Qss, Pss), tdo_u = fma(nre, Qv, Pv), rdo_u = fma(nre, Pv, Qv);     // before the common factor 1/(1 - rinf^2 e^2)
    o.tsd = mul(tsd_u, d.invden);                                     // :656-659
    o.rsd = mul(fma(nre, Pss, Qss), d.invden);
    o.tdo = mul(tdo_u, d.invden);
    o.rdo = mul(rdo_u, d.invden);
    const f2 T3 = mul(fma(rdo_u, Qss, mul(tdo_u, Pss)), d.rinf_invden);
#if PB_F32_ALG
    // T1 + T2 - T3 as nested FMAs (:668-678) and the soil coupling in the A1 A2 form of sail_couple (round 2n):
    //   rsot = rsod + rsos + [A1 A2 / dn + (tsstoo - tss too)] rs,  A1 = tss + tsd, A2 = too + tdo -- every term >= 0, nothing cancels in float
    const f2 Ts = fma(mul(cK, u1), fma(neg(J1ks), F2(too), F2(g[G_Z])), fma(mul(ck, v1), fma(neg(J1ko), F2(tss), F2(g[G_Z])), neg(T3)));
    f2 ss;
    if constexpr (AB) ss = fma(g[G_FSL], rho, mul(tau, g[G_FTL]));    // rsos = w lai sumint, lai sumint folded into Fs, Ft by the prologue   :607, :706
    else ss = mul(fma(g[G_FS], rho, mul(tau, g[G_FT])), g[G_LAISUM]);
    o.rso = fma(Ts, d.inv_omr2, ss);                                  // rsos + rsod                              :678, :706, :710
    const f2 A1 = fma(tsd_u, d.invden, F2(tss)), A2 = fma(tdo_u, d.invden, F2(too));
    o.rsot = fma(fma(mul(A1, A2), d.invdn, F2(g[G_DTT])), rs, o.rso);
    if (need & (1u << O_RDDT)) o.rddt = fma(mul(d.tdd, d.tdd), d.rs_invdn, d.rdd);
    if (need & (1u << O_RSDT)) o.rsdt = fma(mul(A1, d.tdd), d.rs_invdn, o.rsd);
    if (need & (1u << O_RDOT)) o.rdot = fma(mul(d.tdd, A2), d.rs_invdn, o.rdo);
#else
    const f2 w = fma(g[G_FS], rho, mul(tau, g[G_FT]));               // :607
    const f2 T1 = mul(mul(cK, u1), fma(neg(J1ks), F2(too), F2(g[G_Z])));   // :668-675
    const f2 T2 = mul(mul(ck, v1), fma(neg(J1ko), F2(tss), F2(g[G_Z])));
    const f2 rsod = mul(sub(add(T1, T2), T3), d.inv_omr2);            // :678
    o.rso = fma(g[G_LAISUM], w, rsod);                                // :706, :710
    // soil coupling                                                                        :721-726
    const f2 rsodt = mul(fma(add(o.tsd, tss), o.tdo, mul(fma(tss, d.rsrdd, o.tsd), too)), d.rs_invdn);
    o.rsot = add(fma(g[G_TSSTOO], rs, o.rso), rsodt);
    if (need & (1u << O_RDDT)) o.rddt = fma(mul(d.tdd, d.tdd), d.rs_invdn, d.rdd);
    if (need & (1u << O_RSDT)) o.rsdt = fma(mul(add(o.tsd, tss), d.tdd), d.rs_invdn, o.rsd);
    if (need & (1u << O_RDOT)) o.rdot = fma(mul(d.tdd, add(o.tdo, too)), d.rs_invdn, o.rdo);
#endif
}

// ---------------------------------------------------------------------------------------------
// band-mean partial sums: segmented warp reduction over the windows that touch this tile
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mean_partials_inl(const WindowMap* __restrict__ win, const double* __restrict__ srf, T v, int tile, int lane,
                                                  T* __restrict__ dst /* [ntotal] */, bool f32) {
    if (!is_f32<T>() && f32) v = (T)(float)v;
    const int ns = win->nslots[tile];
    for (int j = 0; j < ns; ++j) {
        const int slot = win->slot[tile][j];
        T x = ((win->mask[tile][j] >> lane) & 1u) ? v : T(0);
        if (srf) x *= (T)srf[slot * 32 + lane];                  // spectral-response weighting (zero outside the window)
        const T sum = warp_sum(x);
        if (lane == 0) dst[slot] = sum;
    }
}
// The specialised band-mean instances (compile-time mean mask, plain boxcar windows): the window segments of the thread's NU
// tiles are reduced NU at a time -- slot j of every tile in interleaved butterflies, so that the shuffle latency of one hides
// behind the others.  Unused slot entries of the map have an empty mask (their sum is discarded).
// What a warp needs of the window map is the same for every item it processes (its tile group is fixed), so it is read ONCE, before
// the item loop: per tile of the thread, `member` = bit j set when this lane's band lies inside segment j of the tile, `myslot` =
// the slot id of segment `lane` (lanes 0 .. nslots-1; after the xor butterfly every lane holds the sum, so lane j stores segment j),
// `nmax` = most segments of any of the NU tiles.  No global-memory reads of the map inside the item loop.
template <int NU>
struct MeanPlan {
    unsigned member[NU];   // unpacked map: bit j = this lane lies inside segment j of the tile
    int myslot[NU];        // unpacked map: slot id of segment `lane`; packed map: slot id of the piece that ENDS at this lane, else -1
    unsigned same[NU];     // packed map: bit k = lane - 2^k belongs to the same piece as this lane
    int nmax;              // unpacked map: most segments of any of the NU tiles
};
template <int NU>
__device__ __forceinline__ void mean_plan_init(const WindowMap* __restrict__ win, int tile0, int lane, MeanPlan<NU>& p) {
    p.nmax = 0;
    if (win->packed) {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int mine = win->piece[(tile0 + u) * 32 + lane];
            unsigned sm = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int other = __shfl_up_sync(0xffffffffu, mine, 1 << k);
                sm |= (lane >= (1 << k) && other == mine && mine >= 0) ? (1u << k) : 0u;
            }
            const int next = __shfl_down_sync(0xffffffffu, mine, 1);
            p.same[u] = sm;
            p.myslot[u] = (mine >= 0 && (lane == 31 || next != mine)) ? mine : -1;
            p.member[u] = 0;
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int ns = win->nslots[tile0 + u];
        unsigned m = 0;
        for (int j = 0; j < ns; ++j) m |= ((win->mask[tile0 + u][j] >> lane) & 1u) << j;
        p.member[u] = m;
        p.myslot[u] = lane < ns ? win->slot[tile0 + u][lane] : -1;
        p.same[u] = 0;
        p.nmax = max(p.nmax, ns);
    }
}
template <int NU, typename T>
__device__ __forceinline__ void mean_partialsN_inl(const MeanPlan<NU>& p, bool packed, const T (&vin)[NU], int lane, T* __restrict__ dst, bool f32) {
    T v[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) v[u] = (!is_f32<T>() && f32) ? (T)(float)vin[u] : vin[u];
    if (packed) {
        // segmented inclusive scan over the disjoint pieces of each tile: after the five steps the last lane of a piece holds its sum
#ifndef PB_MEANS_NOREDUCE      /* timing experiment only: what the kernel costs WITHOUT the reduction (results are wrong) */
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const T y = __shfl_up_sync(0xffffffffu, v[u], 1 << k);
                if ((p.same[u] >> k) & 1u) v[u] += y;
            }
        }
#endif
#pragma unroll
        for (int u = 0; u < NU; ++u)
            if (p.myslot[u] >= 0) dst[p.myslot[u]] = v[u];
        return;
    }
    for (int j = 0; j < p.nmax; ++j) {
        T x[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) x[u] = ((p.member[u] >> j) & 1u) ? v[u] : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < NU; ++u) x[u] += __shfl_xor_sync(0xffffffffu, x[u], o);
        }
        if (lane == j) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (p.myslot[u] >= 0) dst[p.myslot[u]] = x[u];
        }
    }
}
template <typename T>
__device__ __noinline__ void mean_partials(const WindowMap* __restrict__ win, const double* __restrict__ srf, T v, int tile, int lane,
                                           T* __restrict__ dst, bool f32) {
    mean_partials_inl(win, srf, v, tile, lane, dst, f32);
}

enum { MODE_PROSAIL = 0, MODE_SAIL = 1, MODE_PROSPECT = 2, MODE_SOIL = 3 };

// Template value of PLANES meaning "planes and means are selected at run time from SailArgs" (the general
// drop-in path).  Any other value is a compile-time plane mask (and MEANS a compile-time mean mask): the LUT
// kernels, where per-item pointer tests, predicated stores and call overhead would cost issue slots.
constexpr unsigned RUNTIME = 0x80000000u;

template <typename T, int NU>
struct EmitCtx {
    const SailArgs<T>& a;
    bool valid[NU];      // band u of this thread exists (< 2101: only the last tile group has lanes past the end); recomputed per chunk from band0
    int tile0;           // 32-band tile of band 0 (band u is in tile0 + u): index into the window map
    int lane, nt;
    long long orow;      // item * pitch + band[0]   (band[u] = band[0] + 32 u)
    T* pdst;             // partial sums of this item
    int mi;              // running index of the mean plane
    MeanPlan<NU> plan;   // this warp's share of the window map (specialised band-mean instances)
};

template <unsigned PLANES, unsigned MEANS, int PLANE, typename T, int NU>
__device__ __forceinline__ void emit(EmitCtx<T, NU>& e, const T (&v)[NU]) {
    if (PLANES == RUNTIME) {
        if (e.a.out[PLANE]) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (e.valid[u]) __stcs(e.a.out[PLANE] + e.orow + 32 * u, v[u]);
        }
        if (e.a.mean_mask & (1u << PLANE)) {
#pragma unroll
            for (int u = 0; u < NU; ++u) mean_partials(e.a.win, e.a.srf, v[u], e.tile0 + u, e.lane, e.pdst + e.mi * e.nt, e.a.f32_means);
            ++e.mi;
        }
    } else {
        if ((PLANES >> PLANE) & 1u) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (e.valid[u]) __stcs(e.a.out[PLANE] + e.orow + 32 * u, v[u]);
        }
        if ((MEANS >> PLANE) & 1u) {       // specialised instances: never launched with spectral-response weights (dispatch_bands)
            if constexpr (!is_f32<T>()) {      // double: NU interleaved butterflies (+6%); float: the per-tile loop is faster (issue-bound)
                mean_partialsN_inl<NU>(e.plan, e.a.packed != 0, v, e.lane, e.pdst + e.mi * e.nt, e.a.f32_means);
            } else {
#pragma unroll
                for (int u = 0; u < NU; ++u) mean_partials_inl(e.a.win, (const double*)nullptr, v[u], e.tile0 + u, e.lane, e.pdst + e.mi * e.nt, e.a.f32_means);
            }
            ++e.mi;
        }
    }
}
// the same value for every band of the thread (bare-soil constants)
template <unsigned PLANES, unsigned MEANS, int PLANE, typename T, int NU>
__device__ __forceinline__ void emit_const(EmitCtx<T, NU>& e, T c) {
    T v[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) v[u] = c;
    emit<PLANES, MEANS, PLANE>(e, v);
}

// resident CTAs per SM the kernel is compiled for: double needs 65536 / threads registers (one CTA fills the register
// file); float fits in half of that, so two CTAs share an SM and cover the MUFU / LDS latencies with twice the warps
template <typename T> __host__ __device__ constexpr int min_ctas() { return std::is_same<T, float>::value ? PB_F32_CTAS : PB_MIN_CTAS; }

// per-plane values of all bands of a thread, gathered from the per-band structs
#define PB_GATHER(dst, arr, field)                       \
    T dst[NU];                                           \
    _Pragma("unroll") for (int u = 0; u < NU; ++u) dst[u] = arr[u] field

// ---------------------------------------------------------------------------------------------
// main kernel.  1-D grid of independent warps; a tile group = NU x 32 consecutive bands, a chunk = CHUNK items.  Two walks over the
// (tile group, chunk) pairs, chosen per instance by the host (launch_bands_g):
//   * strided (nwarps_main = R x ngroups, R = resident warps / ngroups rounded down): warp gw owns the tile group gw % ngroups and
//     walks the chunks gw / ngroups, + R, + 2 R, ... -- all groups sweep the items together, so a row of an output plane is written
//     by all of its groups at about the same time (DRAM write locality: contiguous shares cost the leaf-only and four-plane
//     instances 7-11 %, profiles/r2n_variants*.txt).  resident % ngroups warps stay idle (16 of 1776 for 22 groups).
//   * contiguous (nwarps_main = 0; the one-plane instances, and launches with fewer pairs than resident warps): warp gw takes the
//     share [U gw / nwarps, U (gw + 1) / nwarps) of the U = ngroups x nchunks pairs in group-major order; a share may span several
//     groups, the per-group state (band constants in shared memory, window plan) is rebuilt when the group changes.  Every
//     resident warp works (+1.0 % for the fused fp64 BRF instance).
// The CTA only shares the read-only math tables.
// ---------------------------------------------------------------------------------------------
// G1 = true: one geometry per parameter set (the LUT shape) -- every item starts a new parameter set, so nothing
// is carried between loop iterations and the register allocator sees purely loop-local state.
// CAN = false: the batch is PROSPECT-5 (no anthocyanin term: one FP64 instruction per band and item less in the leaf stage)
template <typename T, typename CF, int MODE, unsigned PLANES, unsigned MEANS, bool G1, bool CAN = true>
__global__ void __launch_bounds__(CF::THREADS, min_ctas<T>()) k_bands(const SailArgs<T> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename Tables<T>::type MT;
    constexpr int NU = CF::NU;
    SailSmem<T, CF, MODE>& sm = *reinterpret_cast<SailSmem<T, CF, MODE>*>(smem_raw);
    const T zero_ = T(0), one = T(1);
    // the warp index through a shuffle: tells the compiler it is warp-uniform (uniform registers / uniform addressing of
    // the per-warp staging buffers and of everything derived from the tile group)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int gw = blockIdx.x * CF::WARPS + warp;                     // global warp id
    Stage<T>* const stage = sm.stage[warp];
    uint64_t* const bar = sm.bar[warp];

    {   // math tables -> shared memory
        const uint4* src = reinterpret_cast<const uint4*>(a.tables);
        uint4* dst = reinterpret_cast<uint4*>(&sm.mt);
        for (int i = threadIdx.x; i < (int)(sizeof(MT) / 16); i += CF::THREADS) dst[i] = src[i];
    }
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();       // the only CTA-wide barrier: tables and barriers are ready

    const long long nchunks = (a.nitems + CHUNK - 1) / CHUNK;
    // this warp's walk over (group slot, chunk) pairs: `remaining` of them, starting at (tslot, chunk); the next chunk is chunk + step,
    // wrapping from wrap_hi to wrap_lo with a change of group
    long long remaining = 0, chunk = 0, step = 1, wrap_lo = 0, wrap_hi = nchunks;
    int tslot = 0;
    if (gw < a.nwarps_main) {
        const int R = a.nwarps_main / a.ngroups;
        const int range = gw / a.ngroups;
        tslot = gw - range * a.ngroups;
        chunk = range; step = R; wrap_hi = a.nchunks_main;
        remaining = chunk < wrap_hi ? (wrap_hi - chunk + R - 1) / R : 0;
    } else if (gw < a.nwarps) {
        // share index: with a full grid the shares of a CTA are taken a grid apart (warp w of CTA b: share w gridDim + b), so that every
        // SM works on (almost) all tile groups at once -- the groups differ in cost (leaf_kall_all skips pigment terms from 784 /
        // 1100 nm on), and neighbouring shares on one SM would leave whole SMs finishing early or late
        const int sh = a.nwarps == (int)gridDim.x * CF::WARPS && a.nwarps_main == 0 ? warp * (int)gridDim.x + (int)blockIdx.x : gw;
        const long long e = sh - a.nwarps_main, E = a.nwarps - a.nwarps_main, nx = nchunks - a.nchunks_main;
        const long long nunits = (long long)a.ngroups * nx;
        const long long t0 = nunits * e / E, t1 = nunits * (e + 1) / E;
        tslot = (int)(t0 / nx);
        wrap_lo = a.nchunks_main;
        chunk = wrap_lo + (t0 - tslot * nx);
        remaining = t1 - t0;
        // (Shares weighted by the cost of a group -- 1000 : 984 : 973 for the three kz classes -- measured no better than equal
        //  counts with this spread: 24.98 vs 24.84 ms, profiles/r2n_variants_kz.txt.)
    }
    if (lane == 0 && remaining > 0) stage_issue(a, &stage[0], &bar[0], chunk);

    int grp = 0, band0 = lane;                     // tile group and this thread's bands (packed map: lanes): band0 + 32 u
    int kz = 0;                                    // which pigment tables vanish over the whole tile group (leaf_kall_all)
    BandConst<T> c[NU];
    // fused fp64 instances keep the interface constants in shared memory (register pressure); everything else in registers
    constexpr bool CSMEM = consts_in_smem<T, MODE>();
    constexpr int KCOL = CF::KCOL;
    const T* kc = sm.kc + threadIdx.x * KCOL;      // this thread's 2 NU adjacent values of every plane pair of the once-per-item constants
    // fp32 mode with two bands per thread: the arithmetic of the item loop runs in packed FP32 (FFMA2 ...), both bands per instruction
    constexpr bool PACKED = std::is_same<T, float>::value && NU == 2;
    constexpr bool AB = MODE == MODE_PROSAIL && !std::is_same<T, float>::value;   // Ap / Am form of the leaf-dependent factors (sail_diffuse)
    BandConst<f2> c2;
    MeanPlan<NU> plan{};
    // per-group state: the thread's band constants (shared-memory columns / registers) and its share of the window map
    auto enter_group = [&](int slot) {
    grp = a.only_active_tiles ? a.win->active_group[NU - 2][slot] : slot;
    band0 = grp * (32 * NU) + lane;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int lane_u = band0 + 32 * u;                            // NBP = 66 * 32: always in range
        const T* bc = a.bandc + (a.packed ? a.win->band[lane_u] : lane_u);   // packed map: gather through the lane -> wavelength list
        T* k = sm.kc + threadIdx.x * KCOL + u;                        // only this thread reads these back: no barrier needed
        k[CF::koff(KC_KAB)] = bc[C_KAB * NBP]; k[CF::koff(KC_KXC)] = bc[C_KXC * NBP]; k[CF::koff(KC_KAN)] = bc[C_KAN * NBP];
        k[CF::koff(KC_KBR)] = bc[C_KBR * NBP]; k[CF::koff(KC_KW)] = bc[C_KW * NBP]; k[CF::koff(KC_KM)] = bc[C_KM * NBP];
        k[CF::koff(KC_RS1)] = bc[C_RS1 * NBP]; k[CF::koff(KC_RS2)] = bc[C_RS2 * NBP];
        c[u] = load_band_const(bc);
        if constexpr (CSMEM) {                               // the fast path reads the folded terms from shared memory; c[] is dead in the loop
            k[CF::koff(KC_TA21)] = c[u].ta21; k[CF::koff(KC_RALF)] = c[u].ralf; k[CF::koff(KC_T1221)] = c[u].t1221;
            k[CF::koff(KC_R12)] = c[u].r12; k[CF::koff(KC_R21)] = c[u].r21;
        }
    }
    if constexpr (PACKED) {
        c2.talf = F2(c[0].talf, c[1].talf); c2.ralf = F2(c[0].ralf, c[1].ralf); c2.t12 = F2(c[0].t12, c[1].t12);
        c2.r12 = F2(c[0].r12, c[1].r12); c2.t21 = F2(c[0].t21, c[1].t21); c2.r21 = F2(c[0].r21, c[1].r21);
    }
    if constexpr (PLANES != RUNTIME && MEANS != 0 && !is_f32<T>()) mean_plan_init<NU>(a.win, NU * grp, lane, plan);
    if constexpr (MODE == MODE_PROSAIL || MODE == MODE_PROSPECT) {
        bool z0 = true, z1 = true;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const T* k = sm.kc + threadIdx.x * KCOL + u;
            z0 = z0 && k[CF::koff(KC_KAB)] == T(0) && k[CF::koff(KC_KXC)] == T(0);
            z1 = z1 && k[CF::koff(KC_KAN)] == T(0) && k[CF::koff(KC_KBR)] == T(0);
        }
        z0 = __all_sync(0xffffffffu, z0);
        z1 = __all_sync(0xffffffffu, z1);
        kz = z0 ? (z1 ? 2 : 1) : 0;
    }
    };
    unsigned need;
    if (PLANES == RUNTIME) {
        need = a.mean_mask;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) need |= a.out[i] ? (1u << i) : 0u;
    } else {
        need = PLANES | MEANS;
    }
    constexpr bool ANY_MEANS = PLANES == RUNTIME || MEANS != 0;
    const int nt = ANY_MEANS ? a.win->ntotal : 0;
    const int nmean = ANY_MEANS ? a.nmean : 0;

    bool new_group = true;
    for (unsigned it = 0; remaining > 0; --remaining, ++it) {
        if (new_group) {
            __syncwarp();
            enter_group(tslot);
            new_group = false;
        }
        const int buf = it & 1;
        long long next = chunk + step;                                // the chunk after this one: the first of the next group at the end of a group
        if (next >= wrap_hi) { next = wrap_lo; new_group = true; }
        if (lane == 0 && remaining > 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            stage_issue(a, &stage[buf ^ 1], &bar[buf ^ 1], next);
        }
        mbar_wait(&bar[buf], (it >> 1) & 1u);

        {
            const long long i0 = chunk * CHUNK;
            const int n = (int)min((long long)CHUNK, a.nitems - i0);
            const long long s0 = G1 ? i0 : (long long)((unsigned)i0 / (unsigned)a.G);
            long long s = s0;
            int g = G1 ? 0 : (int)(i0 - s0 * a.G);
            bool fresh = true;
            T rho[NU], tau[NU], rs[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) rho[u] = tau[u] = rs[u] = zero_;
            Diffuse<T> d[NU];
            Diffuse<f2> dp;          // packed twin of d[] (fp32 mode)
            int thr_hi = 0;
            bool nocanopy = false;
            const T* sr = stage[buf].s;
            const T* gr = stage[buf].g;
            EmitCtx<T, NU> e{a, {}, NU * grp, lane, nt, i0 * a.out_pitch + band0, ANY_MEANS ? a.partial + i0 * nmean * nt : nullptr, 0, plan};
#pragma unroll
            for (int u = 0; u < NU; ++u)      // (the packed map stores no planes and reads no spectra: the flag is not consulted there)
                e.valid[u] = band0 + 32 * u < NB;
            // software pipelining of the double LUT instances (PB_PIPELINE): tau(kall) of item j + 1 is evaluated inside the big FP64 block
            // of item j (after the plate / Stokes stage; in the fused kernels it shares a basic block with the canopy stage).  Always
            // evaluated, also past the last item of the chunk: the record read then is whatever follows in the staging buffer
            // (in bounds of the Stage struct; tau_fast is total on any bit pattern) and the value is dropped.
            constexpr bool PIPE = PB_PIPELINE && G1 && !is_f32<T>() && PLANES != RUNTIME && (MODE == MODE_PROSAIL || MODE == MODE_PROSPECT);
            T tau_n[NU];
            unsigned rare_n = 0;
            if constexpr (PIPE) leaf_tau<CF, CAN>(kc, sr, kz, sm.mt, tau_n, rare_n);
            for (int j = 0; j < n; ++j) {
                e.mi = 0;
                if (MODE == MODE_SOIL) {
                    soil_rho_all<CF>(kc, sr, rs);
                    emit<PLANES, MEANS, O_RHO>(e, rs);
                } else if (MODE == MODE_PROSPECT) {
                    if constexpr (PACKED) {
                        f2 r2, t2;
                        leaf_optics_x2<CF, CAN>(c2, kc, sr, kz, sm.mt, r2, t2);
                        rho[0] = r2.x; rho[1] = r2.y; tau[0] = t2.x; tau[1] = t2.y;
                    } else if constexpr (PIPE) {
                        T tau_c[NU];
#pragma unroll
                        for (int u = 0; u < NU; ++u) tau_c[u] = tau_n[u];
                        leaf_tau_repair<CF, CAN>(kc, sr, kz, tau_c, rare_n);
                        leaf_tau<CF, CAN>(kc, sr + REC, kz, sm.mt, tau_n, rare_n);          // next item: independent work next to the Stokes stage
                        leaf_layers_all<CF, CSMEM>(c, kc, sr, sm.mt, tau_c, rho, tau);
                    } else {
                        leaf_optics<CF, CSMEM, CAN>(c, kc, sr, kz, sm.mt, rho, tau);
                    }
                    emit<PLANES, MEANS, O_LEAF_R>(e, rho);
                    emit<PLANES, MEANS, O_LEAF_T>(e, tau);
                    if (PLANES == RUNTIME && (need & (7u << O_LEAF_KA))) {       // absorptance, extinction, single-scattering albedo
                        T ka[NU], ke[NU], om[NU];
#pragma unroll
                        for (int u = 0; u < NU; ++u) {                           // models.py:1066-1068
                            ka[u] = one - rho[u] - tau[u];
                            ke[u] = rho[u] + ka[u];
                            om[u] = rho[u] / ke[u];
                        }
                        emit<PLANES, MEANS, O_LEAF_KA>(e, ka);
                        emit<PLANES, MEANS, O_LEAF_KE>(e, ke);
                        emit<PLANES, MEANS, O_LEAF_OM>(e, om);
                    }
                    if (PLANES == RUNTIME && (need & (31u << O_LAYER_R))) {      // PROSPECT.r, t, Ra, Ta, denom (models.py:959)
                        T q[NU][5];
#pragma unroll
                        for (int u = 0; u < NU; ++u) {
                            const BandConst<T> cc = c[u];     // a copy: the callee takes a reference, and c[] itself must stay in registers
                            leaf_one_layer(cc, leaf_kall<CF>(kc + u, sr), q[u]);
                        }
                        { PB_GATHER(v, q, [0]); emit<PLANES, MEANS, O_LAYER_R>(e, v); }
                        { PB_GATHER(v, q, [1]); emit<PLANES, MEANS, O_LAYER_T>(e, v); }
                        { PB_GATHER(v, q, [2]); emit<PLANES, MEANS, O_LAYER_RA>(e, v); }
                        { PB_GATHER(v, q, [3]); emit<PLANES, MEANS, O_LAYER_TA>(e, v); }
                        { PB_GATHER(v, q, [4]); emit<PLANES, MEANS, O_LAYER_DENOM>(e, v); }
                    }
                } else {
                    if (G1 || fresh) {
                        if (MODE == MODE_PROSAIL) {
                            if constexpr (PACKED) {
                                f2 r2, t2;
                                leaf_optics_x2<CF, CAN>(c2, kc, sr, kz, sm.mt, r2, t2);
                                rho[0] = r2.x; rho[1] = r2.y; tau[0] = t2.x; tau[1] = t2.y;
                            } else if constexpr (PIPE) {
                                T tau_c[NU];
#pragma unroll
                                for (int u = 0; u < NU; ++u) tau_c[u] = tau_n[u];
                                leaf_tau_repair<CF, CAN>(kc, sr, kz, tau_c, rare_n);
                                leaf_layers_all<CF, CSMEM>(c, kc, sr, sm.mt, tau_c, rho, tau);
                                leaf_tau<CF, CAN>(kc, sr + REC, kz, sm.mt, tau_n, rare_n);  // next item: independent work next to the canopy stage
                            } else {
                                leaf_optics<CF, CSMEM, CAN>(c, kc, sr, kz, sm.mt, rho, tau);
                            }
                            soil_rho_all<CF>(kc, sr, rs);
                        } else {
#pragma unroll
                            for (int u = 0; u < NU; ++u) {
                                const int bb = e.valid[u] ? band0 + 32 * u : 0;
                                rho[u] = a.in_r[s * a.in_pitch_r + bb];
                                tau[u] = a.in_t[s * a.in_pitch_t + bb];
                                rs[u] = a.in_rho[s * a.in_pitch_rho + bb];
                            }
                        }
                        if constexpr (PACKED) {
                            sail_diffuse_f32x2<MODE == MODE_PROSAIL>(F2(rho[0], rho[1]), F2(tau[0], tau[1]), F2(rs[0], rs[1]), sr, dp);
                        } else {
#pragma unroll
                            for (int u = 0; u < NU; ++u) sail_diffuse<MODE == MODE_PROSAIL>(rho[u], tau[u], rs[u], sr, sm.mt, d[u]);
                        }
                        // |k - m| lai <= 1e-3  <=>  |k - m| <= 1e-3 / lai   (threshold on the high word)
                        thr_hi = hi_word(sr[S_THR]);
                        nocanopy = hi_word(sr[S_NEGLAI]) == 0;                   // bare soil: the record carries +0.0 there (any canopy: a negative number); an integer test, once per set
                        fresh = false;
                    }
                    {
                        // The canopy stage runs for EVERY item, bare soil (lai <= 0, models.py:609-634) included: with the record of
                        // such an item (-lai -> 0, tss = too = tsstoo = 1, z = lai sumint = 0, written by the prologue) every formula
                        // stays finite and already yields rho_surface to rounding; the exact values of the reference are patched in
                        // afterwards under a (rare, warp-uniform) branch.  Round 1 branched BEFORE the stage: the compiler hoisted
                        // directional-stage work above that branch and spilled five doubles per item across it (STL/LDL in the
                        // item loop of the 168-register instance), and the loop body lost one of its large basic blocks.
                        Canopy<T> o[NU];
                        if constexpr (PACKED) {
                            Canopy<f2> op;
                            op.rddt = op.rsdt = op.rdot = F2(0.0f);
                            sail_directional_f32x2<MODE == MODE_PROSAIL && PB_F32_ALG>(F2(rho[0], rho[1]), F2(tau[0], tau[1]), F2(rs[0], rs[1]), dp, sr, gr, need, op);
                            o[0] = Canopy<T>{op.rsot.x, op.rddt.x, op.rsdt.x, op.rdot.x, op.rso.x, op.rsd.x, op.tsd.x, op.rdo.x, op.tdo.x};
                            o[1] = Canopy<T>{op.rsot.y, op.rddt.y, op.rsdt.y, op.rdot.y, op.rso.y, op.rsd.y, op.tsd.y, op.rdo.y, op.tdo.y};
                        } else {
                            bool rare[NU];
                            bool any_rare = false;
#pragma unroll
                            for (int u = 0; u < NU; ++u) {
                                o[u].rddt = o[u].rsdt = o[u].rdot = zero_;
                                rare[u] = sail_directional_fast<AB>(rho[u], tau[u], rs[u], d[u], sr, gr, thr_hi, need, o[u]);
                                any_rare |= rare[u];
                            }
                            if constexpr (!std::is_same<T, float>::value) {
                                if (__builtin_expect(any_rare, 0)) {
#pragma unroll
                                    for (int u = 0; u < NU; ++u)
                                        if (rare[u]) sail_directional<AB>(rho[u], tau[u], rs[u], d[u], sr, gr, thr_hi, need, o[u]);
                                }
                            }
                        }
                        if constexpr (PACKED) {      // the diffuse planes come from the packed state
                            d[0].rdd = dp.rdd.x; d[1].rdd = dp.rdd.y; d[0].tdd = dp.tdd.x; d[1].tdd = dp.tdd.y;
                        }
                        if (__builtin_expect(nocanopy, 0)) {                     // lai <= 0: exactly models.py:609-634
#pragma unroll
                            for (int u = 0; u < NU; ++u) {
                                o[u].rsot = o[u].rddt = o[u].rsdt = o[u].rdot = rs[u];
                                o[u].rso = o[u].rsd = o[u].tsd = o[u].rdo = o[u].tdo = zero_;
                                d[u].rdd = zero_; d[u].tdd = one;
                            }
                        }
                        { PB_GATHER(v, o, .rsot); emit<PLANES, MEANS, O_RSOT>(e, v); }
                        { PB_GATHER(v, o, .rddt); emit<PLANES, MEANS, O_RDDT>(e, v); }
                        { PB_GATHER(v, o, .rsdt); emit<PLANES, MEANS, O_RSDT>(e, v); }
                        { PB_GATHER(v, o, .rdot); emit<PLANES, MEANS, O_RDOT>(e, v); }
                        { PB_GATHER(v, o, .rso); emit<PLANES, MEANS, O_RSO>(e, v); }
                        { PB_GATHER(v, d, .rdd); emit<PLANES, MEANS, O_RDD>(e, v); }
                        { PB_GATHER(v, d, .tdd); emit<PLANES, MEANS, O_TDD>(e, v); }
                        { PB_GATHER(v, o, .rsd); emit<PLANES, MEANS, O_RSD>(e, v); }
                        { PB_GATHER(v, o, .tsd); emit<PLANES, MEANS, O_TSD>(e, v); }
                        { PB_GATHER(v, o, .rdo); emit<PLANES, MEANS, O_RDO>(e, v); }
                        { PB_GATHER(v, o, .tdo); emit<PLANES, MEANS, O_TDO>(e, v); }
                    }
                    if constexpr (PACKED) { d[0].m = dp.m.x; d[1].m = dp.m.y; }
                    { PB_GATHER(v, d, .m); emit<PLANES, MEANS, O_KE>(e, v); }
                    emit<PLANES, MEANS, O_LEAF_R>(e, rho);
                    emit<PLANES, MEANS, O_LEAF_T>(e, tau);
                    emit<PLANES, MEANS, O_RHO>(e, rs);
                }
                gr += REC;
                e.orow += a.out_pitch;
                if (ANY_MEANS) e.pdst += nmean * nt;
                if (G1) { ++s; sr += REC; }
                else if (++g == a.G) { g = 0; ++s; sr += REC; fresh = true; }
            }
        }
        __syncwarp();      // all lanes are done with stage[buf] before lane 0 refills it in the next iteration
        chunk = next;
        tslot += new_group ? 1 : 0;
    }
}
#undef PB_GATHER

// band means = ordered sum of a window's slot partials / band count (or / sum of the response weights)
template <typename T>
__global__ void k_finalize_means(const WindowMap* __restrict__ win, const T* __restrict__ partial, long long nrows /* nitems*nmean */,
                                 T* __restrict__ means, float* __restrict__ means_f32) {
    const int nw = win->nwindows, nt = win->ntotal;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * nw) return;
    const long long row = idx / nw;
    const int w = (int)(idx - row * nw);
    double acc = 0.0;
    for (int sl = win->wstart[w]; sl < win->wend[w]; ++sl) acc += (double)partial[row * nt + sl];
    const double mean = acc / win->wsum[w];
    if (means) means[idx] = (T)mean;
    if (means_f32) means_f32[idx] = (float)mean;
}

// ---------------------------------------------------------------------------------------------
// prologue: one warp per (parameter set, geometry).  Lanes 0..17 are the 18 leaf-inclination bins
// (LIDF + volume-scattering sums, warp-shuffle reductions); lanes 0..19 are the 20 hotspot steps.
// ---------------------------------------------------------------------------------------------
struct PrologueArgs {
    long long nsamples;
    int G;                 // geometries per parameter set
    int geom_per_sample;   // 1: iza/vza/raa have nsamples entries (G must be 1); 0: G entries shared by all sets
    int lidf_type;         // 0 campbell (models.py:283-332), 1 verhoef (models.py:335-392)
    int has_leaf, has_soil, has_canopy;
    const double *N, *Cab, *Cxc, *Cbr, *Cw, *Cm, *Can;
    const double *soil_refl, *soil_moist;
    const double *lai, *hot, *lidf_a, *lidf_b;
    const double *iza, *vza, *raa;   // radians, already normalised like core/_core.py:136-146
    void* srec;            // [nsamples][REC] of T (the kernel's template type: double, or float for the fp32 mode)
    void* grec;            // [nitems][REC] of T
    double* geom_out;      // optional [nitems][NGEOM]: VolScatt ks, ko, bf, Fs, Ft; tss, too, tsstoo; chi_s, chi_o, frho, ftau of the last leaf bin
    double* lidf_out;      // optional [nsamples][18]
};

#define PB_PI 3.14159265358979323846

// volume scattering functions for one leaf zenith angle (degrees) -- models.py:177-258
__device__ __forceinline__ void volscatt_volume(double tts, double tto, double psi, double ttl_deg, double& chi_s, double& chi_o,
                                                double& frho, double& ftau) {
    const double cts = cos(tts), cto = cos(tto), sts = sin(tts), sto = sin(tto), cospsi = cos(psi);
    const double tl = ttl_deg * (PB_PI / 180.0);
    const double clza = cos(tl), slza = sin(tl);
    const double cs = clza * cts, co = clza * cto, ss = slza * sts, so = slza * sto;
    double cosbts = 5.0, cosbto = 5.0;
    if (fabs(ss) > 1e-6) cosbts = -cs / ss;
    if (fabs(so) > 1e-6) cosbto = -co / so;
    double bts, ds, bto, do_;
    if (fabs(cosbts) < 1.0) { bts = acos(cosbts); ds = ss; } else { bts = PB_PI; ds = cs; }
    chi_s = 2.0 / PB_PI * ((bts - PB_PI * 0.5) * cs + sin(bts) * ss);
    if (fabs(cosbto) < 1.0) { bto = acos(cosbto); do_ = so; }
    else if (tto < 90.0 * PB_PI / 180.0) { bto = PB_PI; do_ = co; }
    else { bto = 0.0; do_ = -co; }
    chi_o = 2.0 / PB_PI * ((bto - PB_PI * 0.5) * co + sin(bto) * so);
    const double btran1 = fabs(bts - bto);
    const double btran2 = PB_PI - fabs(bts + bto - PB_PI);
    double bt1, bt2, bt3;
    if (psi <= btran1) { bt1 = psi; bt2 = btran1; bt3 = btran2; }
    else {
        bt1 = btran1;
        if (psi <= btran2) { bt2 = psi; bt3 = btran2; } else { bt2 = btran2; bt3 = psi; }
    }
    const double t1 = 2.0 * cs * co + ss * so * cospsi;
    double t2 = 0.0;
    if (bt2 > 0.0) t2 = sin(bt2) * (2.0 * ds * do_ + ss * so * cos(bt1) * cos(bt3));
    const double denom = 2.0 * PB_PI * PB_PI;
    frho = fmax(((PB_PI - bt2) * t1 + t2) / denom, 0.0);
    ftau = fmax((-bt2 * t1 + t2) / denom, 0.0);
}

// LIDF.campbell / LIDF.verhoef / VolScatt.coef for an arbitrary number of leaf-inclination classes `ne` (the reference's
// `n_elements` argument, models.py:82-175, 283-392; SAIL itself always uses 18, which is what the prologue kernels are built
// for).  One thread per parameter set: lidf [n][ne] (always written), geom (optional) [n][G][NGEOM] with ks, ko, bf, Fs, Ft in
// slots 0-4 and chi_s, chi_o, frho, ftau of the last class in slots 8-11 (the rest zero).  Plain libm arithmetic in the
// reference's order; not on the hot path.
__global__ void __launch_bounds__(128) k_volscatt_general(long long n, int G, int lidf_type, const double* __restrict__ la_,
                                                          const double* __restrict__ lb_, const double* __restrict__ iza,
                                                          const double* __restrict__ vza, const double* __restrict__ raa, int ne,
                                                          double* __restrict__ geom, double* __restrict__ lidf) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double la = la_[s], lb = lb_ ? lb_[s] : 0.0;
    double* L = lidf + s * ne;
    const double step = 90.0 / ne;
    if (lidf_type == 0) {                                             // Campbell, models.py:299-330
        const double excent = exp(-1.6184e-5 * la * la * la + 2.1145e-3 * la * la - 1.2390e-1 * la + 3.2491);
        const double e2 = excent * excent;
        const double alph = excent / sqrt(fabs(1.0 - e2)), alph2 = alph * alph;
        double sum0 = 0.0, Fprev = 0.0;
        for (int i = 0; i <= ne; ++i) {                               // F at the ne + 1 class edges; freq_i = |F_i - F_{i+1}|
            const double tl = (i * step) * (PB_PI / 180.0);
            double F;
            if (excent == 1.0) {
                F = cos(tl);
            } else {
                const double tn = tan(tl);
                const double x = excent / sqrt(1.0 + e2 * tn * tn), x2 = x * x;
                if (excent > 1.0) { const double p = sqrt(alph2 + x2); F = x * p + alph2 * log(x + p); }
                else { const double q = sqrt(alph2 - x2); F = x * q + alph2 * asin(x / alph); }
            }
            if (i > 0) { L[i - 1] = fabs(Fprev - F); sum0 += L[i - 1]; }
            Fprev = F;
        }
        for (int i = 0; i < ne; ++i) L[i] = L[i] / sum0;
    } else {                                                          // Verhoef, models.py:366-390: from the last class edge down to 0
        double Fhi = 1.0;
        for (int i = ne - 1; i >= 0; --i) {
            const double tl1 = (i * step) * (PB_PI / 180.0);
            double f;
            if (la > 1.0) {
                f = 1.0 - cos(tl1);
            } else {
                double x = 2.0 * tl1, y = 0.0, dx;
                const double p = x;
                int itn = 0;
                do {
                    y = la * sin(x) + 0.5 * lb * sin(2.0 * x);
                    dx = 0.5 * (y - x + p);
                    x += dx;
                } while (fabs(dx) >= 1e-8 && ++itn < 4096);
                f = (2.0 * y + p) / PB_PI;
            }
            L[i] = Fhi - f;
            Fhi = f;
        }
    }
    if (!geom) return;
    for (int g = 0; g < G; ++g) {                                     // VolScatt.coef, models.py:144-175
        const double tts = iza[g], tto = vza[g], psi = raa[g];
        const double cts = cos(tts), cto = cos(tto);
        double ks = 0.0, ko = 0.0, bf = 0.0, Fs = 0.0, Ft = 0.0, chi_s = 0.0, chi_o = 0.0, frho = 0.0, ftau = 0.0;
        for (int i = 0; i < ne; ++i) {
            const double ttl = i * step + step * 0.5;
            const double cttl = cos(ttl * (PB_PI / 180.0));
            volscatt_volume(tts, tto, psi, ttl, chi_s, chi_o, frho, ftau);
            ks += chi_s / cts * L[i];
            ko += chi_o / cto * L[i];
            bf += cttl * cttl * L[i];
            Fs += frho * PB_PI / (cts * cto) * L[i];
            Ft += ftau * PB_PI / (cts * cto) * L[i];
        }
        double* o = geom + (s * G + g) * NGEOM;
        o[0] = ks; o[1] = ko; o[2] = bf; o[3] = Fs; o[4] = Ft; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
        o[8] = chi_s; o[9] = chi_o; o[10] = frho; o[11] = ftau;
    }
}

// VolScatt.volume(lza) for n geometries and one leaf zenith angle (degrees): out [n][4] = chi_s, chi_o, frho, ftau (models.py:177-258)
__global__ void k_volume(int n, const double* __restrict__ iza, const double* __restrict__ vza, const double* __restrict__ raa, double lza_deg,
                         double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double chi_s, chi_o, frho, ftau;
    volscatt_volume(iza[i], vza[i], raa[i], lza_deg, chi_s, chi_o, frho, ftau);
    out[4 * i] = chi_s; out[4 * i + 1] = chi_o; out[4 * i + 2] = frho; out[4 * i + 3] = ftau;
}

// S_CWX, S_CBRX of a leaf record: see the S_ enum and leaf_kall_all
__device__ __forceinline__ void leaf_poison_copies(double (&srec)[REC]) {
    const bool ok01 = isfinite(srec[S_CAB]) && isfinite(srec[S_CXC]);
    srec[S_CBRX] = ok01 ? srec[S_CBR] : nan("");
    srec[S_CWX] = ok01 && isfinite(srec[S_CAN]) && isfinite(srec[S_CBR]) ? srec[S_CW] : nan("");
}

// S_NEGLAI of a canopy: -lai; in double records (PB_EXPS) pre-multiplied with 2048 log2(e) for exp_neg_scaled, |value| < 2^30
template <typename T>
__device__ __forceinline__ double neglai_rec(double lai) {
#if PB_EXPS
    if constexpr (!std::is_same<T, float>::value) return -fmin(lai * PB_EXPS_SCALE, 1.0e9);
#endif
    return -lai;
}

// write one 16-element record as T (double: 128 B, float: 64 B) with 16-byte stores
template <typename T>
__device__ __forceinline__ void store_rec(void* base, long long idx, const double (&rec)[REC]) {
    if constexpr (std::is_same<T, float>::value) {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(base) + idx * REC);
#pragma unroll
        for (int i = 0; i < REC / 4; ++i) dst[i] = make_float4((float)rec[4 * i], (float)rec[4 * i + 1], (float)rec[4 * i + 2], (float)rec[4 * i + 3]);
    } else {
        double2* dst = reinterpret_cast<double2*>(static_cast<double*>(base) + idx * REC);
#pragma unroll
        for (int i = 0; i < REC / 2; ++i) dst[i] = make_double2(rec[2 * i], rec[2 * i + 1]);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) k_prologue(const PrologueArgs a) {
    const int lane = threadIdx.x & 31;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nitems = a.nsamples * a.G;
    if (item >= nitems) return;
    const long long s = item / a.G;
    const int g = (int)(item - s * a.G);

    // ---- sample record (written by lane 0 of the g == 0 warp) ----
    double srec[REC];
#pragma unroll
    for (int i = 0; i < REC; ++i) srec[i] = 0.0;
    if (a.has_leaf) {
        const double N = a.N[s];
        srec[S_CAB] = a.Cab[s] / N; srec[S_CXC] = a.Cxc[s] / N; srec[S_CAN] = a.Can ? a.Can[s] / N : 0.0;
        srec[S_CBR] = a.Cbr[s] / N; srec[S_CW] = a.Cw[s] / N; srec[S_CM] = a.Cm[s] / N;
        srec[S_NM1] = N - 1.0;
        srec[S_LMAX] = 500.0 / fabs(N - 1.0);     // clamp of log2 b in b^(N-1) (pow_pos_scaled); inf for N = 1
        leaf_poison_copies(srec);
    }
    if (a.has_soil) {
        const double sr = a.soil_refl[s], mo = a.soil_moist[s];
        srec[S_SOIL1] = sr * mo;                  // rho = sRef (moisture rsoil1 + (1-moisture) rsoil2), models.py:1917
        srec[S_SOIL2] = sr * (1.0 - mo);
    }
    if (!a.has_canopy) {
        if (g == 0 && lane == 0) store_rec<T>(a.srec, s, srec);
        return;
    }

    const double lai = a.lai[s], hot = a.hot[s];
    const double la = a.lidf_a[s], lb = a.lidf_b ? a.lidf_b[s] : 0.0;
    const long long gi = a.geom_per_sample ? s : g;
    const double tts = a.iza[gi], tto = a.vza[gi], psi = a.raa[gi];

    // ---- LIDF over 18 bins of 5 degrees ----
    double lidf = 0.0;
    if (a.lidf_type == 0) {                       // Campbell, models.py:301-330
        const double excent = exp(-1.6184e-5 * la * la * la + 2.1145e-3 * la * la - 1.2390e-1 * la + 3.2491);
        double freq = 0.0;
        if (lane < 18) {
            const double tl1 = (lane * 5.0) * PB_PI / 180.0, tl2 = ((lane + 1.0) * 5.0) * PB_PI / 180.0;
            const double tn1 = tan(tl1), tn2 = tan(tl2);
            const double x1 = excent / sqrt(1.0 + excent * excent * tn1 * tn1);
            const double x2 = excent / sqrt(1.0 + excent * excent * tn2 * tn2);
            if (excent == 1.0) {
                freq = fabs(cos(tl1) - cos(tl2));
            } else {
                const double alph = excent / sqrt(fabs(1.0 - excent * excent));
                const double alph2 = alph * alph, x12 = x1 * x1, x22 = x2 * x2;
                if (excent > 1.0) {
                    const double p1 = sqrt(alph2 + x12), p2 = sqrt(alph2 + x22);
                    freq = fabs(x1 * p1 + alph2 * log(x1 + p1) - (x2 * p2 + alph2 * log(x2 + p2)));
                } else {
                    const double q1 = sqrt(alph2 - x12), q2 = sqrt(alph2 - x22);
                    freq = fabs(x1 * q1 + alph2 * asin(x1 / alph) - (x2 * q2 + alph2 * asin(x2 / alph)));
                }
            }
        }
        const double sum0 = warp_sum(freq);
        lidf = freq / sum0;
    } else {                                      // Verhoef, models.py:366-390
        // F(theta) at theta = lane*5 deg; lidf_i = F((i+1)*5) - F(i*5), F(90) = 1.
        double f = 1.0;
        if (lane < 18) {
            const double tl1 = (lane * 5.0) * (PB_PI / 180.0);
            if (la > 1.0) {
                f = 1.0 - cos(tl1);
            } else {
                double x = 2.0 * tl1, y = 0.0, dx;
                const double p = x;
                int itn = 0;
                do {
                    y = la * sin(x) + 0.5 * lb * sin(2.0 * x);
                    dx = 0.5 * (y - x + p);
                    x += dx;
                } while (fabs(dx) >= 1e-8 && ++itn < 4096);
                f = (2.0 * y + p) / PB_PI;
            }
        }
        const double fnext = __shfl_down_sync(0xffffffffu, f, 1);    // lane 17 receives lane 18's f = 1.0
        lidf = lane < 18 ? fnext - f : 0.0;
    }
    if (a.lidf_out && g == 0 && lane < 18) a.lidf_out[s * 18 + lane] = lidf;

    // ---- LIDF-weighted volume-scattering sums, models.py:144-175 ----
    double ksli = 0.0, koli = 0.0, bfli = 0.0, sobli = 0.0, sofli = 0.0;
    const double cts = cos(tts), cto = cos(tto);
    double last[4] = {0.0, 0.0, 0.0, 0.0};        // chi_s, chi_o, frho, ftau of this lane's bin; lane 17 = the last bin (models.py:159)
    if (lane < 18) {
        const double ttl = lane * 5.0 + 2.5;
        double chi_s, chi_o, frho, ftau;
        volscatt_volume(tts, tto, psi, ttl, chi_s, chi_o, frho, ftau);
        const double cttl = cos(ttl * (PB_PI / 180.0));
        ksli = chi_s / cts * lidf;
        koli = chi_o / cto * lidf;
        bfli = cttl * cttl * lidf;
        sobli = frho * PB_PI / (cts * cto) * lidf;
        sofli = ftau * PB_PI / (cts * cto) * lidf;
        last[0] = chi_s; last[1] = chi_o; last[2] = frho; last[3] = ftau;
    }
    if (a.geom_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i) last[i] = __shfl_sync(0xffffffffu, last[i], 17);
    }
    const double k = warp_sum(ksli), K = warp_sum(koli), bf = warp_sum(bfli), Fs = warp_sum(sobli), Ft = warp_sum(sofli);

    // ---- direct transmittances and the hotspot integral, models.py:664-666, 687-702, 731-763 ----
    const bool nocanopy = !(lai > 0.0);
    double tss = 1.0, too = 1.0, tsstoo = 1.0, sumint = 0.0, z = 0.0, alf = 1e36;
    if (!nocanopy) {
        tss = exp(-k * lai);
        too = exp(-K * lai);
        z = (1.0 - exp(-(k + K) * lai)) / (k + K);                   // Jfunc2(ks, ko, lai)
        const double tants = tan(tts), tanto = tan(tto);
        const double dso = sqrt(tants * tants + tanto * tanto - 2.0 * tants * tanto * cos(psi));
        if (hot > 0.0) alf = (dso / hot) * 2.0 / (k + K);
        if (alf == 0.0) {                                             // the pure hotspot
            tsstoo = tss;
            sumint = (1.0 - tss) / (k * lai);
        } else {
            const double fhot = lai * sqrt(K * k);
            const double fint = (1.0 - exp(-alf)) * 0.05;
            // lane i holds step istep = i + 1; (x1, y1, f1) come from lane i-1
            double x2 = 1.0, y2 = 0.0, f2 = 1.0;
            if (lane < 20) {
                if (lane < 19) x2 = -log(1.0 - (lane + 1) * fint) / alf;
                y2 = -(K + k) * lai * x2 + fhot * (1.0 - exp(-alf * x2)) / alf;
                f2 = exp(y2);
            }
            double x1 = __shfl_up_sync(0xffffffffu, x2, 1), y1 = __shfl_up_sync(0xffffffffu, y2, 1),
                   f1 = __shfl_up_sync(0xffffffffu, f2, 1);
            if (lane == 0) { x1 = 0.0; y1 = 0.0; f1 = 1.0; }
            const double term = lane < 20 ? (f2 - f1) * (x2 - x1) / (y2 - y1) : 0.0;
            sumint = warp_sum(term);
            tsstoo = __shfl_sync(0xffffffffu, f2, 19);
            if (isnan(sumint)) sumint = 0.0;
        }
    }

    if (lane == 0) {
        double r[REC];
        r[G_K] = k; r[G_KO] = K;
        r[G_SDB] = 0.5 * (k + bf); r[G_SDF] = 0.5 * (k - bf);        // models.py:583-588
        r[G_DOB] = 0.5 * (K + bf); r[G_DOF] = 0.5 * (K - bf);
        r[G_FS] = Fs; r[G_FT] = Ft;
        r[G_TSS] = tss; r[G_TOO] = too; r[G_TSSTOO] = tsstoo;
        r[G_LAISUM] = lai * sumint; r[G_Z] = z; r[G_FSL] = Fs * (lai * sumint); r[G_FTL] = Ft * (lai * sumint); r[G_DTT] = tsstoo - tss * too;
        store_rec<T>(a.grec, item, r);
        if (g == 0) {
            if (!a.has_leaf) { srec[S_DDB] = 0.5 * (1.0 + bf); srec[S_DDF] = 0.5 * (1.0 - bf); }      // with a leaf: S_CWX, S_CBRX (set above)
            srec[S_NEGLAI] = nocanopy ? 0.0 : neglai_rec<T>(lai); srec[S_LAI] = lai;   // bare soil (+0.0 IS the no-canopy flag): the band kernel's canopy stage runs with lai = 0
            srec[S_THR] = nocanopy ? 0.0 : 1e-3 / lai;
            srec[S_BF] = bf;
            store_rec<T>(a.srec, s, srec);
        }
        if (a.geom_out) {
            double* o = a.geom_out + item * NGEOM;
            o[0] = k; o[1] = K; o[2] = bf; o[3] = Fs; o[4] = Ft; o[5] = tss; o[6] = too; o[7] = tsstoo;
            o[8] = last[0]; o[9] = last[1]; o[10] = last[2]; o[11] = last[3];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// prologue, throughput variant: one THREAD per (parameter set, geometry), used for large batches.
// Everything that is uniform for an item (cos/sin of the sun/view angles, exp terms) is computed once per
// item instead of once per lane, the 18 leaf bins and 20 hotspot steps are loops in the same order as the
// reference (so the sums round like models.py:154-175 and 749-759), and the trigonometry of
// VolScatt.volume is reduced to two acos per bin: with cb = cos(beta), sb = sqrt(1 - cb^2),
//   cos(btran1) = cos(bts - bto),  sin(btran1) = |sin(bts - bto)|,
//   cos(btran2) = cos(bts + bto),  sin(btran2) = |sin(bts + bto)|      (btran2 = pi - |bts + bto - pi|).
// ---------------------------------------------------------------------------------------------
struct AngleTables {            // filled on the host by prosail_create with the reference's own expressions
    double cl[18], sl[18];      // cos / sin of the bin centres 2.5, 7.5, ... deg   (np.radians(ttl), models.py:202-203)
    double tan_e[19], cos_e[19];   // tan / cos of the bin edges i * 5 deg, i = 0..18  (rad(i * step), models.py:306-311)
    double edge[19];            // the edges in radians as LIDF.verhoef computes them (np.radians(angle), models.py:372)
};

// One leaf-inclination bin of VolScatt.volume / coef (models.py:154-173, 177-258) for one geometry:
// b = { chi_s/cos(tts), chi_o/cos(tto), frho pi/(cos cos), ftau pi/(cos cos), chi_s, chi_o, frho, ftau }.
// The trigonometry is reduced to two acos per bin with angle-addition identities (see k_prologue_t).
__device__ __forceinline__ void volscatt_bin(const AngleTables* __restrict__ at, int i, double tto, double sts, double cts, double sto,
                                             double cto, double spsi, double cpsi, double psi, double (&b)[8]) {
    const double ctscto = cts * cto;
    const bool view_up = tto < 90.0 * PB_PI / 180.0;
    const double clza = at->cl[i], slza = at->sl[i];
    const double cs = clza * cts, co = clza * cto, ss = slza * sts, so = slza * sto;
    double cbs = 5.0, cbo = 5.0;
    if (fabs(ss) > 1e-6) cbs = -cs / ss;
    if (fabs(so) > 1e-6) cbo = -co / so;
    double bts, ds, sbs, bto, do_, sbo;
    if (fabs(cbs) < 1.0) { bts = acos(cbs); ds = ss; sbs = sqrt(1.0 - cbs * cbs); }
    else { bts = PB_PI; ds = cs; cbs = -1.0; sbs = 0.0; }
    const double chi_s = 2.0 / PB_PI * ((bts - PB_PI * 0.5) * cs + sbs * ss);
    if (fabs(cbo) < 1.0) { bto = acos(cbo); do_ = so; sbo = sqrt(1.0 - cbo * cbo); }
    else if (view_up) { bto = PB_PI; do_ = co; cbo = -1.0; sbo = 0.0; }
    else { bto = 0.0; do_ = -co; cbo = 1.0; sbo = 0.0; }
    const double chi_o = 2.0 / PB_PI * ((bto - PB_PI * 0.5) * co + sbo * so);
    const double btran1 = fabs(bts - bto);
    const double btran2 = PB_PI - fabs(bts + bto - PB_PI);
    const double cc = cbs * cbo, sscp = sbs * sbo, sc = sbs * cbo, cs_ = cbs * sbo;
    const double cos1 = cc + sscp, sin1 = fabs(sc - cs_);        // btran1
    const double cos2 = cc - sscp, sin2 = fabs(sc + cs_);        // btran2
    double bt2, sin_bt2, cos_bt1, cos_bt3;
    if (psi <= btran1) { bt2 = btran1; sin_bt2 = sin1; cos_bt1 = cpsi; cos_bt3 = cos2; }
    else if (psi <= btran2) { bt2 = psi; sin_bt2 = spsi; cos_bt1 = cos1; cos_bt3 = cos2; }
    else { bt2 = btran2; sin_bt2 = sin2; cos_bt1 = cos1; cos_bt3 = cpsi; }
    const double t1 = 2.0 * cs * co + ss * so * cpsi;
    double t2 = 0.0;
    if (bt2 > 0.0) t2 = sin_bt2 * (2.0 * ds * do_ + ss * so * cos_bt1 * cos_bt3);
    const double denom = 2.0 * PB_PI * PB_PI;
    const double frho = fmax(((PB_PI - bt2) * t1 + t2) / denom, 0.0);
    const double ftau = fmax((-bt2 * t1 + t2) / denom, 0.0);
    b[0] = chi_s / cts; b[1] = chi_o / cto; b[2] = frho * PB_PI / ctscto; b[3] = ftau * PB_PI / ctscto;
    b[4] = chi_s; b[5] = chi_o; b[6] = frho; b[7] = ftau;
}

// The bin terms depend on the geometry only: when a batch shares its geometry list (LUT shapes), they are computed once per
// (geometry, bin) into a table [G][18][8] and every (parameter set, geometry) item just weights them with its LIDF.
__global__ void __launch_bounds__(128) k_geom_bins(int G, const double* __restrict__ iza, const double* __restrict__ vza,
                                                   const double* __restrict__ raa, const AngleTables* __restrict__ at, double* __restrict__ ws) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * 18) return;
    const int g = t / 18, i = t - g * 18;
    double sts, cts, sto, cto, spsi, cpsi;
    sincos(iza[g], &sts, &cts);
    sincos(vza[g], &sto, &cto);
    sincos(raa[g], &spsi, &cpsi);
    double b[8];
    volscatt_bin(at, i, vza[g], sts, cts, sto, cto, spsi, cpsi, raa[g], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) ws[(long long)t * 8 + j] = b[j];
}

// LIDF over 18 bins of 5 degrees for one parameter set (thread-sequential form; bins in the reference's order)
__device__ __forceinline__ void lidf_bins(int lidf_type, double la, double lb, const AngleTables* __restrict__ at, double (&lidf)[18]) {
    if (lidf_type == 0) {                       // Campbell, models.py:301-330
        const double excent = exp(-1.6184e-5 * la * la * la + 2.1145e-3 * la * la - 1.2390e-1 * la + 3.2491);
        const double e2 = excent * excent;
        double sum0 = 0.0;
        if (excent == 1.0) {
#pragma unroll 1
            for (int i = 0; i < 18; ++i) { lidf[i] = fabs(at->cos_e[i] - at->cos_e[i + 1]); sum0 += lidf[i]; }
        } else {
            const double alph = excent / sqrt(fabs(1.0 - e2));
            const double alph2 = alph * alph;
            const bool prolate = excent > 1.0;
            double Fprev = 0.0;
#pragma unroll 1
            for (int i = 0; i <= 18; ++i) {       // F at the 19 bin edges; freq_i = |F_i - F_{i+1}|
                const double tn = at->tan_e[i];
                const double x = excent / sqrt(1.0 + e2 * tn * tn);
                const double x2 = x * x;
                double F;
                if (prolate) { const double p = sqrt(alph2 + x2); F = x * p + alph2 * log(x + p); }
                else { const double q = sqrt(alph2 - x2); F = x * q + alph2 * asin(x / alph); }
                if (i > 0) { lidf[i - 1] = fabs(Fprev - F); sum0 += lidf[i - 1]; }
                Fprev = F;
            }
        }
#pragma unroll 1
        for (int i = 0; i < 18; ++i) lidf[i] = lidf[i] / sum0;
    } else {                                      // Verhoef, models.py:366-390: F(theta) from 85 deg down to 0
        double Fhi = 1.0;
#pragma unroll 1
        for (int i = 17; i >= 0; --i) {
            const double tl1 = at->edge[i];
            double f;
            if (la > 1.0) {
                f = 1.0 - at->cos_e[i];
            } else {
                double x = 2.0 * tl1, y = 0.0, dx;
                const double p = x;
                int itn = 0;
                do {
                    double sx, cx;
                    sincos(x, &sx, &cx);
                    y = la * sx + 0.5 * lb * (2.0 * sx * cx);          // sin 2x = 2 sin x cos x
                    dx = 0.5 * (y - x + p);
                    x += dx;
                } while (fabs(dx) >= 1e-8 && ++itn < 4096);
                f = (2.0 * y + p) / PB_PI;
            }
            lidf[i] = Fhi - f;
            Fhi = f;
        }
    }
}

// LIDF once per parameter set into a workspace [nsamples][18]: used when several geometries share a set, so that the
// Verhoef fixed-point solves (the dominant cost of the prologue) are not repeated per geometry
__global__ void __launch_bounds__(128) k_lidf(const PrologueArgs a, const AngleTables* __restrict__ at, double* __restrict__ ws) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.nsamples) return;
    double lidf[18];
    lidf_bins(a.lidf_type, a.lidf_a[s], a.lidf_b ? a.lidf_b[s] : 0.0, at, lidf);
#pragma unroll 1
    for (int i = 0; i < 18; ++i) ws[s * 18 + i] = lidf[i];
}

template <typename T, bool LIDF_WS, bool GEOM_WS>
__global__ void __launch_bounds__(128) k_prologue_t(const PrologueArgs a, const AngleTables* __restrict__ at,
                                                    const double* __restrict__ lidf_ws, const double* __restrict__ geom_ws) {
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nitems = a.nsamples * a.G;
    if (item >= nitems) return;
    const long long s = item / a.G;
    const int g = (int)(item - s * a.G);

    double srec[REC];
#pragma unroll
    for (int i = 0; i < REC; ++i) srec[i] = 0.0;
    if (g == 0) {
        if (a.has_leaf) {
            const double N = a.N[s];
            srec[S_CAB] = a.Cab[s] / N; srec[S_CXC] = a.Cxc[s] / N; srec[S_CAN] = a.Can ? a.Can[s] / N : 0.0;
            srec[S_CBR] = a.Cbr[s] / N; srec[S_CW] = a.Cw[s] / N; srec[S_CM] = a.Cm[s] / N;
            srec[S_NM1] = N - 1.0;
            srec[S_LMAX] = 500.0 / fabs(N - 1.0);     // clamp of log2 b in b^(N-1) (pow_pos_scaled); inf for N = 1
            leaf_poison_copies(srec);
        }
        if (a.has_soil) {
            const double sr = a.soil_refl[s], mo = a.soil_moist[s];
            srec[S_SOIL1] = sr * mo;              // rho = sRef (moisture rsoil1 + (1-moisture) rsoil2), models.py:1917
            srec[S_SOIL2] = sr * (1.0 - mo);
        }
    }
    if (!a.has_canopy) {
        if (g == 0) store_rec<T>(a.srec, s, srec);
        return;
    }

    const double lai = a.lai[s], hot = a.hot[s];
    const double la = a.lidf_a[s], lb = a.lidf_b ? a.lidf_b[s] : 0.0;
    const long long gi = a.geom_per_sample ? s : g;
    const double tts = a.iza[gi], tto = a.vza[gi], psi = a.raa[gi];
    double sts, cts, sto, cto, spsi, cpsi;
    sincos(tts, &sts, &cts);
    sincos(tto, &sto, &cto);
    sincos(psi, &spsi, &cpsi);

    // ---- LIDF over 18 bins of 5 degrees: computed here, or read back when k_lidf ran first (several geometries per set) ----
    double lidf[18];
    if (LIDF_WS) {
#pragma unroll 1
        for (int i = 0; i < 18; ++i) lidf[i] = lidf_ws[s * 18 + i];
    } else {
        lidf_bins(a.lidf_type, la, lb, at, lidf);
    }
    if (a.lidf_out && g == 0) {
#pragma unroll 1
        for (int i = 0; i < 18; ++i) a.lidf_out[s * 18 + i] = lidf[i];
    }

    // ---- LIDF-weighted volume-scattering sums, models.py:144-258 ----
    double k = 0.0, K = 0.0, bf = 0.0, Fs = 0.0, Ft = 0.0;
    double last_chi_s = 0.0, last_chi_o = 0.0, last_frho = 0.0, last_ftau = 0.0;
#pragma unroll 1
    for (int i = 0; i < 18; ++i) {
        double b[8];                                  // ws wo wr wt chi_s chi_o frho ftau of this bin
        if (GEOM_WS) {
            const double4* row = reinterpret_cast<const double4*>(geom_ws + ((long long)gi * 18 + i) * 8);
            const double4 r0 = row[0];
            b[0] = r0.x; b[1] = r0.y; b[2] = r0.z; b[3] = r0.w;
            if (i == 17) { const double4 r1 = row[1]; b[4] = r1.x; b[5] = r1.y; b[6] = r1.z; b[7] = r1.w; }
        } else {
            volscatt_bin(at, i, tto, sts, cts, sto, cto, spsi, cpsi, psi, b);
        }
        const double clza = at->cl[i];
        const double w = lidf[i];
        if (i == 17) { last_chi_s = b[4]; last_chi_o = b[5]; last_frho = b[6]; last_ftau = b[7]; }
        k += b[0] * w;
        K += b[1] * w;
        bf += clza * clza * w;
        Fs += b[2] * w;
        Ft += b[3] * w;
    }

    // ---- direct transmittances and the hotspot integral, models.py:664-666, 687-702, 731-763 ----
    const bool nocanopy = !(lai > 0.0);
    double tss = 1.0, too = 1.0, tsstoo = 1.0, sumint = 0.0, z = 0.0, alf = 1e36;
    if (!nocanopy) {
        tss = exp(-k * lai);
        too = exp(-K * lai);
        z = (1.0 - tss * too) / (k + K);                              // Jfunc2(ks, ko, lai): exp(-(k+K) lai) = tss too
        const double tants = sts / cts, tanto = sto / cto;
        const double dso = sqrt(fmax(tants * tants + tanto * tanto - 2.0 * tants * tanto * cpsi, 0.0));
        if (hot > 0.0) alf = (dso / hot) * 2.0 / (k + K);
        if (alf == 0.0) {                                             // the pure hotspot
            tsstoo = tss;
            sumint = (1.0 - tss) / (k * lai);
        } else {
            const double fhot = lai * sqrt(K * k);
            const double fint = (1.0 - exp(-alf)) * 0.05;
            double x1 = 0.0, y1 = 0.0, f1 = 1.0;
#pragma unroll 1
            for (int istep = 1; istep <= 20; ++istep) {
                const double x2 = istep < 20 ? -log(1.0 - istep * fint) / alf : 1.0;
                const double y2 = -(K + k) * lai * x2 + fhot * (1.0 - exp(-alf * x2)) / alf;
                const double f2 = exp(y2);
                sumint += (f2 - f1) * (x2 - x1) / (y2 - y1);
                x1 = x2; y1 = y2; f1 = f2;
            }
            tsstoo = f1;
            if (isnan(sumint)) sumint = 0.0;
        }
    }

    double grec[REC];
    grec[G_K] = k; grec[G_KO] = K;
    grec[G_SDB] = 0.5 * (k + bf); grec[G_SDF] = 0.5 * (k - bf);      // models.py:583-588
    grec[G_DOB] = 0.5 * (K + bf); grec[G_DOF] = 0.5 * (K - bf);
    grec[G_FS] = Fs; grec[G_FT] = Ft;
    grec[G_TSS] = tss; grec[G_TOO] = too; grec[G_TSSTOO] = tsstoo;
    grec[G_LAISUM] = lai * sumint; grec[G_Z] = z; grec[G_FSL] = Fs * (lai * sumint); grec[G_FTL] = Ft * (lai * sumint); grec[G_DTT] = tsstoo - tss * too;
    store_rec<T>(a.grec, item, grec);
    if (g == 0) {
        if (!a.has_leaf) { srec[S_DDB] = 0.5 * (1.0 + bf); srec[S_DDF] = 0.5 * (1.0 - bf); }          // with a leaf: S_CWX, S_CBRX (set above)
        srec[S_NEGLAI] = nocanopy ? 0.0 : neglai_rec<T>(lai); srec[S_LAI] = lai;   // bare soil (+0.0 IS the no-canopy flag): the band kernel's canopy stage runs with lai = 0
        srec[S_THR] = nocanopy ? 0.0 : 1e-3 / lai;
        srec[S_BF] = bf;
        store_rec<T>(a.srec, s, srec);
    }
    if (a.geom_out) {
        double* o = a.geom_out + item * NGEOM;
        o[0] = k; o[1] = K; o[2] = bf; o[3] = Fs; o[4] = Ft; o[5] = tss; o[6] = too; o[7] = tsstoo;
        o[8] = last_chi_s; o[9] = last_chi_o; o[10] = last_frho; o[11] = last_ftau;      // as left behind by models.py:159
    }
}

// ---------------------------------------------------------------------------------------------
// test / calibration kernels
// ---------------------------------------------------------------------------------------------
// op: 0 rcp, 1 sqrt_pos, 2 exp_neg, 3 log2_pos, 4 exp2_any, 5 tau_plate, 6 tau_full, 7 rcp_q, 8 sqrt_q_raw (x > 0),
//     9 exp_neg_scaled (exp(x), x <= 0, as m = -x / 8 and the record value c = -8 * 2048 log2(e)), 10 pow_pos_scaled (x^0.75, x >= 1)
__global__ void k_test_math(int op, long long n, const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ tables) {
    extern __shared__ __align__(16) unsigned char test_smem[];              // sizeof(MathTables) bytes of dynamic shared memory
    MathTables& mt = *reinterpret_cast<MathTables*>(test_smem);
    for (int i = threadIdx.x; i < (int)(sizeof(MathTables) / 8); i += blockDim.x) reinterpret_cast<double*>(&mt)[i] = tables[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = in[i];
    double y = 0.0;
    switch (op) {
        case 0: y = rcp(x); break;
        case 1: y = sqrt_pos(x); break;
        case 2: y = exp_neg(x, mt.exp2t); break;
        case 3: y = log2_pos(x, mt.logt); break;
        case 4: y = exp2_any(x, mt.exp2t); break;
        case 6: y = tau_full(x); break;
        case 7: y = rcp_q(x); break;
        case 8: y = sqrt_q_raw(x); break;
        case 9: y = exp_neg_scaled(x * -0.125, -8.0 * PB_EXPS_SCALE, mt.exp2t); break;
        case 10: y = pow_pos_scaled(x, 0.75, __double2hiint(500.0 / 0.75), mt); break;
        default: y = tau_plate(x, mt.tau); break;
    }
    out[i] = y;
}

// DFMA throughput microbenchmark: 8 independent chains per thread, `iters` x 8 x 4 FMAs each.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

}  // namespace pb
