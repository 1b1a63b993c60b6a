# Builds libprosail_b200.so (sm_100a only) in-tree.  `make` here or __graft_entry__.build().
NVCC ?= nvcc
# --register-usage-level=3: ptxas' default (5) costs the multi-geometry instances 4% and the fp32 / four-plane ones 1% (profiles/r1g_variants.txt)
NVFLAGS ?= -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas --register-usage-level=3 -Xcompiler -fPIC
# tuning experiments: make variant NAME=nu3 EXTRA=-DPB_NU=3  ->  pyrism_b200/lib/variants/libprosail_nu3.so (PYRISM_B200_LIB selects it)
EXTRA ?=
SRC := pyrism_b200/csrc/prosail_api.cu
HDR := pyrism_b200/csrc/prosail_kernels.cuh pyrism_b200/csrc/math64.cuh pyrism_b200/csrc/math32.cuh pyrism_b200/csrc/invert_kernels.cuh pyrism_b200/csrc/invert_tc.cuh pyrism_b200/csrc/math_tables.inc include/prosail_b200.h
LIB := pyrism_b200/lib/libprosail_b200.so

all: $(LIB)

$(LIB): $(SRC) $(HDR) Makefile
	mkdir -p pyrism_b200/lib
	$(NVCC) $(NVFLAGS) $(EXTRA) -shared -o $@ $(SRC)

variant: $(SRC) $(HDR)
	mkdir -p pyrism_b200/lib/variants
	$(NVCC) $(NVFLAGS) $(EXTRA) -shared -o pyrism_b200/lib/variants/libprosail_$(NAME).so $(SRC)

ptxas: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -Xptxas -v -shared -o /tmp/libprosail_b200_v.so $(SRC)

# pipe microbenchmarks (results: profiles/r1b_micro_*.txt): make micro && tools/micro/bin/<name> on the GPU box
micro:
	mkdir -p tools/micro/bin
	for f in dfma_lat issue_mix fp64_flavours fp64_operands conv_mix; do $(NVCC) -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/bin/$$f tools/micro/$$f.cu; done

clean:
	rm -f $(LIB)
.PHONY: all clean ptxas variant micro
