# Builds the C-ABI library without Python (what a Rust build.rs / CI job would call).
#   make            -> accumulation_b200/libaccmsm.so (sm_100a)
#   make oracle     -> oracle/liboracle.so            (test infrastructure)
#   make harness    -> tests/host/as_tests            (C++ parity harness; needs both of the above)
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= /usr/bin/g++   # not $$(CXX): build images export CXX to toolchains without libgomp
CSRC := accumulation_b200/csrc
DEPS := $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.inc) include/accmsm.h

all: accumulation_b200/libaccmsm.so

accumulation_b200/libaccmsm.so: $(DEPS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fopenmp \
	    -shared -diag-suppress 550 -o $@ $(CSRC)/accmsm.cu

oracle:
	$(MAKE) -C oracle

harness: accumulation_b200/libaccmsm.so oracle tests/host/as_tests.cpp accumulation_b200/host/ark_mirror.hpp
	$(HOSTCXX) -O2 -std=c++17 -Wall -o tests/host/as_tests tests/host/as_tests.cpp -Laccumulation_b200 -Loracle \
	    -l:libaccmsm.so -l:liboracle.so -Wl,-rpath,$(CURDIR)/accumulation_b200 -Wl,-rpath,$(CURDIR)/oracle -fopenmp -pthread

clean:
	rm -f accumulation_b200/libaccmsm.so tests/host/as_tests tests/host/libhostshim.so
	$(MAKE) -C oracle clean

.PHONY: all oracle harness clean
