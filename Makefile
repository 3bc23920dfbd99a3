# Builds libvokselis_rt.so (the product: CUDA kernels + C ABI + C++ host mirror) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v
SRC := vokselis_b200/csrc
OBJ := build/obj
LIB := vokselis_b200/libvokselis_rt.so

CU := $(SRC)/raycast.cu $(SRC)/sortlast.cu $(SRC)/volume.cu $(SRC)/present.cu $(SRC)/api.cu
CUO := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU))
HDR := $(SRC)/raycast.cuh $(SRC)/vkrt_device.cuh include/vokselis_rt.h vokselis_b200/host/vokselis.hpp

all: $(LIB) build/headless build/microbench

$(OBJ)/%.o: $(SRC)/%.cu $(HDR)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)

$(OBJ)/camera.o: vokselis_b200/host/camera.cpp $(HDR)
	@mkdir -p $(OBJ)
	$(HOSTCXX) -O2 -std=c++17 -fPIC -fvisibility=hidden -Wall -Wextra -c $< -o $@

$(LIB): $(CUO) $(OBJ)/camera.o
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $^ -cudart shared

build/headless: vokselis_b200/host/headless.cpp $(LIB) $(HDR)
	@mkdir -p build
	$(HOSTCXX) -O2 -std=c++17 -Wall -Wextra $< -o $@ -Lvokselis_b200 -lvokselis_rt -Wl,-rpath,'$$ORIGIN/../vokselis_b200'

build/microbench: bench/microbench.cu
	@mkdir -p build
	$(NVCC) -O3 -std=c++17 $(ARCH) -ccbin $(HOSTCXX) -o $@ $<

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB)

.PHONY: all oracle clean
