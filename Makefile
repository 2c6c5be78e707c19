# Builds the product shared library (C ABI + CUDA kernels for sm_100a) in-tree, and the test oracle.
NVCC ?= nvcc
NVCCFLAGS := $(EXTRA) --threads 4 -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
             -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v
LIB := rna_algos_b200/librna_algos_b200.so
SRC := rna_algos_b200/csrc/rna_abi.cu rna_algos_b200/csrc/fold_fastnum.cu rna_algos_b200/csrc/rna_multi.cpp rna_algos_b200/csrc/rna_queue.cpp
HDR := $(wildcard rna_algos_b200/csrc/*.cuh rna_algos_b200/csrc/*.h include/*.h)

all: $(LIB) oracle cli peaks

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC) -lcudart

oracle:
	$(MAKE) -C oracle

# command-line front ends with the reference's options and output formats (cli/rna_cli.cpp); one binary, three names
CLI := bin/rna_algos_b200
cli: $(CLI) rna_algos_b200/tables_standin/standin_turner.tbl
$(CLI): cli/rna_cli.cpp include/rna_algos_b200.h $(LIB)
	mkdir -p bin
	g++ -std=c++17 -O2 -Wall -o $@ cli/rna_cli.cpp -Lrna_algos_b200 -lrna_algos_b200 -Wl,-rpath,'$$ORIGIN/../rna_algos_b200' -Wl,--allow-shlib-undefined
	ln -sf rna_algos_b200 bin/mccaskill_algo && ln -sf rna_algos_b200 bin/centroid_fold && ln -sf rna_algos_b200 bin/durbin_algo
# (stand-in score tables for --standin-tables; the genuine blobs come from tools/ref_dump)
rna_algos_b200/tables_standin/standin_turner.tbl: rna_algos_b200/tables.py
	python -m rna_algos_b200.tables dump rna_algos_b200/tables_standin

# measured roofline denominators for bench.py (FP32 non-FMA issue, MUFU, shared-memory bandwidth): measurement
# infrastructure, not linked by the product library
peaks: tools/_build/libpeaks.so
tools/_build/libpeaks.so: tools/peak_microbench.cu
	mkdir -p tools/_build
	$(NVCC) -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -shared -Xcompiler -fPIC -o $@ $< -lcudart

clean:
	rm -rf $(LIB) bin rna_algos_b200/tables_standin build tools/_build
	$(MAKE) -C oracle clean
.PHONY: all oracle cli peaks clean
