# Builds the product shared library (C ABI + CUDA kernels for sm_100a) in-tree, and the test oracle.
NVCC ?= nvcc
NVCCFLAGS := -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
             -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v
LIB := rna_algos_b200/librna_algos_b200.so
SRC := rna_algos_b200/csrc/rna_abi.cu
HDR := $(wildcard rna_algos_b200/csrc/*.cuh rna_algos_b200/csrc/*.h include/*.h)

all: $(LIB) oracle

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
