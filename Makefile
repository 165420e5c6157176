# Builds the C-ABI shared library (CUDA, sm_100a) and the CPU oracle.
NVCC ?= nvcc
NVFLAGS = -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -pthread
LIB = lite_b200/libliteb200.so

all: $(LIB) oracle tests/cpp/test_facade tests/cpp/smf_host tests/cpp/gen_host examples

SRC = lite_b200/csrc/capi.cu lite_b200/csrc/host_tess.cpp lite_b200/csrc/host_capi.cpp lite_b200/csrc/facade.cpp
HDR = lite_b200/csrc/gen_face.h lite_b200/csrc/kernels.cuh lite_b200/csrc/philox.h lite_b200/csrc/smf_chain.h lite_b200/csrc/smf_ensemble.cuh lite_b200/csrc/host_tess.h include/lite_b200.h include/lite_b200_host.h include/ttessel.h

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean examples

# C++ tests of the facade (run by tests/test_facade_cpp.py)
tests/cpp/test_facade: tests/cpp/test_facade.cpp include/ttessel.h $(LIB)
	g++ -O1 -std=c++17 -Iinclude -o $@ tests/cpp/test_facade.cpp -Llite_b200 -lliteb200 -Wl,-rpath,'$$ORIGIN/../../lite_b200' -pthread

# host build of the SMF chain source for the oracle replay test (run by tests/test_smf_chain_cpu.py)
tests/cpp/smf_host: tests/cpp/smf_host.cpp lite_b200/csrc/smf_chain.h lite_b200/csrc/philox.h include/lite_b200.h
	g++ -O2 -std=c++17 -o $@ tests/cpp/smf_host.cpp

# host build of the general-face source (faces with holes) for the oracle test (run by tests/test_gen_face_cpu.py)
tests/cpp/gen_host: tests/cpp/gen_host.cpp lite_b200/csrc/gen_face.h include/lite_b200.h
	g++ -O2 -std=c++17 -o $@ tests/cpp/gen_host.cpp

# the reference's drivers against this library (need a GPU to run)
examples: examples/pseudo_lik_inference examples/check_acs_ensemble examples/sim_acs
examples/%: examples/%.cpp include/ttessel.h include/lite_b200.h include/lite_b200_host.h $(LIB)
	g++ -O1 -std=c++17 -Iinclude -o $@ $< -Llite_b200 -lliteb200 -Wl,-rpath,'$$ORIGIN/../lite_b200' -pthread
