# Build of libsarw_b200.so (CUDA kernels + C ABI), the sarw CLI host, and the test oracles.
# sm_100a only. -fmad=false / -ffp-contract=off: accept/reject arithmetic must not contract multiply-adds.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
PKG := polymer-sarw_b200
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -Iinclude -I$(PKG)/csrc
LIB := $(PKG)/lib/libsarw_b200.so
CLI := $(PKG)/bin/sarw

all: lib cli oracle

lib: $(LIB)
$(LIB): $(PKG)/csrc/sarw_kernels.cu $(PKG)/csrc/sarw_abi.cu $(PKG)/csrc/sarw_analysis.cu $(PKG)/csrc/sarw_device.cuh $(PKG)/csrc/sarw_cluster.cuh $(PKG)/csrc/sarw_host.h include/sarw_abi.h
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ $(PKG)/csrc/sarw_kernels.cu $(PKG)/csrc/sarw_abi.cu $(PKG)/csrc/sarw_analysis.cu 2> $(PKG)/lib/ptxas.log || (cat $(PKG)/lib/ptxas.log; exit 1)

# same library with per-step cycle counters in the growth kernels (scripts/prof_fastpath.py; load it with SARW_B200_LIB=...)
prof: $(PKG)/lib/libsarw_prof.so
$(PKG)/lib/libsarw_prof.so: $(PKG)/csrc/sarw_kernels.cu $(PKG)/csrc/sarw_abi.cu $(PKG)/csrc/sarw_analysis.cu $(PKG)/csrc/sarw_device.cuh $(PKG)/csrc/sarw_cluster.cuh $(PKG)/csrc/sarw_host.h include/sarw_abi.h
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -DSARW_PROF -shared -o $@ $(PKG)/csrc/sarw_kernels.cu $(PKG)/csrc/sarw_abi.cu $(PKG)/csrc/sarw_analysis.cu

# A/B variants of the library for measurements: make variant NAME=nofence DEFS="-DSARW_NO_PROXY_FENCE"; load with SARW_B200_LIB=...
variant:
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) $(DEFS) -shared -o $(PKG)/lib/libsarw_$(NAME).so $(PKG)/csrc/sarw_kernels.cu $(PKG)/csrc/sarw_abi.cu $(PKG)/csrc/sarw_analysis.cu

HOSTSRC := $(PKG)/host/input_map.cpp $(PKG)/host/polymer_io.cpp
cli: $(CLI) $(PKG)/bin/sarw_hosttest
$(CLI): $(HOSTSRC) $(PKG)/host/main.cpp $(wildcard $(PKG)/host/*.h) include/sarw_abi.h $(LIB)
	mkdir -p $(PKG)/bin
	$(CXX) -O2 -g -std=c++17 -Wall -ffp-contract=off -Iinclude -I$(PKG)/host -o $@ $(HOSTSRC) $(PKG)/host/main.cpp -L$(PKG)/lib -lsarw_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lpthread
# host layer without the GPU library (CPU tests of parser / topology / exporters)
$(PKG)/bin/sarw_hosttest: $(HOSTSRC) $(PKG)/host/hosttest.cpp $(wildcard $(PKG)/host/*.h)
	mkdir -p $(PKG)/bin
	$(CXX) -O2 -g -std=c++17 -Wall -DSARW_HOSTTEST -I$(PKG)/host -o $@ $(HOSTSRC) $(PKG)/host/hosttest.cpp -lpthread

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(PKG)/lib $(PKG)/bin
	$(MAKE) -C oracle clean
.PHONY: all lib cli oracle prof clean variant
