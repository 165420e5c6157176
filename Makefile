# Builds the C-ABI shared library (CUDA, sm_100a) and the CPU oracle.
NVCC ?= nvcc
NVFLAGS = -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
LIB = lite_b200/libliteb200.so

all: $(LIB) oracle

$(LIB): lite_b200/csrc/capi.cu lite_b200/csrc/kernels.cuh include/lite_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ lite_b200/csrc/capi.cu -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
