# libffvd_b200.so for sm_100a, built in-tree (it travels to the GPU box with gpurun).
# The fused-kernel template instantiations are split into ten objects so that `make -j` builds them in parallel.
NVCC      ?= nvcc
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
SRC       := ffvd_b200/csrc
OBJ       := build/obj
LIB       := ffvd_b200/lib/libffvd_b200.so
KINDS     := 0 1
MODES     := 0 1 2 3 4
INST_OBJS := $(foreach k,$(KINDS),$(foreach m,$(MODES),$(OBJ)/fused_inst_$(k)_$(m).o))
HDRS      := $(wildcard $(SRC)/*.cuh) include/ffvd_b200.h

all: $(LIB)

$(OBJ)/fused_inst_%.o: $(SRC)/fused_inst.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DFFVD_SPLIT_BUILD -DFFVD_INST_KIND=$(word 1,$(subst _, ,$*)) -DFFVD_INST_MODE=$(word 2,$(subst _, ,$*)) -c -o $@ $<

$(OBJ)/capi.o: $(SRC)/capi.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DFFVD_SPLIT_BUILD -c -o $@ $<

$(OBJ)/fused_lookup.o: $(SRC)/fused_lookup.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIB): $(OBJ)/capi.o $(OBJ)/fused_lookup.o $(INST_OBJS)
	@mkdir -p ffvd_b200/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $^

# single-translation-unit development builds (see tools/devrun.sh)
dev:
	$(NVCC) $(NVFLAGS) -shared -DFFVD_DEV_MINIMAL -o ffvd_b200/lib/libffvd_b200_dev.so $(SRC)/capi.cu
dev-timing:
	$(NVCC) $(NVFLAGS) -shared -DFFVD_DEV_MINIMAL -DFFVD_PHASE_TIMING -o ffvd_b200/lib/libffvd_b200_dev.so $(SRC)/capi.cu

# diagnostic build with the index assertions of the fused kernels (compute-sanitizer is closed on the development pool);
# exercised by tools/bounds_check.py and tests/test_gpu_parity.py::test_bounds_checked_build
check:
	$(NVCC) $(NVFLAGS) -shared -DFFVD_DEV_MINIMAL -DFFVD_BOUNDS_CHECK -o ffvd_b200/lib/libffvd_b200_check.so $(SRC)/capi.cu

clean:
	rm -rf build ffvd_b200/lib/*.so

.PHONY: all dev dev-timing check clean
