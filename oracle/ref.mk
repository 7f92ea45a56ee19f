# ORACLE — TEST INFRASTRUCTURE ONLY.
# Builds the REFERENCE'S OWN sources, from where they lie under /root/reference (nothing is copied into the repo),
# against oracle/shim (host Kokkos subset, MKL prototypes bound to the oneMKL inside libtorch_cpu.so):
#   _ref/gmres_perf_test   the reference CLI, unmodified (gmres_perf_test.cpp + gmres.cpp + kernels_mkl.cpp + mmio.c)
#   _ref/libref.so         gmres.cpp + kernels_mkl.cpp + oracle/ref_driver.cpp (C entry point with residual logging)
# The reference's own Makefile is NOT used (it needs Kokkos, the MKL SDK and CUDA 10).  Flags: no -DNDEBUG (the
# reference performs MKL calls inside assert()); -fpermissive because the host build maps the Cuda tag
# onto MKL (shim/types_cuda.hpp), which repeats explicit instantiations.
REF      ?= /root/reference
OUT      := _ref
FARM     := $(OUT)/src
CXX       = g++
CC        = gcc
TORCHLIB := $(shell python -c "import os, torch; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
CXXFLAGS  = -O2 -march=x86-64-v3 -fopenmp -fpermissive -w -std=c++14 -fPIC -I shim -I $(FARM)
LDLIBS    = -L$(TORCHLIB) -ltorch_cpu -lc10 -Wl,-rpath,$(TORCHLIB) -fopenmp
SRCS      = gmres.cpp gmres.hpp Orthogonalization.hpp IterUtil.hpp kernels.hpp types.hpp types_mkl.hpp kernels_mkl.cpp \
            gmres_perf_test.cpp LoadMatrix.hpp mmio.c mmio.h

all: $(OUT)/libref.so $(OUT)/gmres_perf_test

$(FARM)/.linked:
	mkdir -p $(FARM)
	for f in $(SRCS); do ln -sf $(REF)/$$f $(FARM)/$$f; done
	touch $@

$(OUT)/gmres.o: $(FARM)/.linked shim/Kokkos_Core.hpp
	$(CXX) $(CXXFLAGS) -c $(FARM)/gmres.cpp -o $@
$(OUT)/kernels_mkl.o: $(FARM)/.linked shim/Kokkos_Core.hpp shim/mkl.h
	$(CXX) $(CXXFLAGS) -c $(FARM)/kernels_mkl.cpp -o $@
$(OUT)/gmres_perf_test.o: $(FARM)/.linked shim/Kokkos_Core.hpp
	$(CXX) $(CXXFLAGS) -include cstring -include sstream -c $(FARM)/gmres_perf_test.cpp -o $@
$(OUT)/mmio.o: $(FARM)/.linked
	$(CC) -O2 -w -fPIC -c $(FARM)/mmio.c -o $@
$(OUT)/mkl_shim.o: shim/mkl_shim.cpp shim/mkl.h
	$(CXX) $(CXXFLAGS) -c shim/mkl_shim.cpp -o $@
$(OUT)/ref_driver.o: ref_driver.cpp $(FARM)/.linked shim/Kokkos_Core.hpp
	$(CXX) $(CXXFLAGS) -include cstring -include sstream -I $(FARM) -c ref_driver.cpp -o $@

$(OUT)/libref.so: $(OUT)/gmres.o $(OUT)/kernels_mkl.o $(OUT)/mkl_shim.o $(OUT)/ref_driver.o $(OUT)/mmio.o
	$(CXX) -shared -o $@ $^ $(LDLIBS)
$(OUT)/gmres_perf_test: $(OUT)/gmres_perf_test.o $(OUT)/gmres.o $(OUT)/kernels_mkl.o $(OUT)/mkl_shim.o $(OUT)/mmio.o
	$(CXX) -o $@ $^ $(LDLIBS)

.PHONY: all

# ---- `b200`: the reference's unmodified CLI + driver on the B200 backend ------------------------------------------
# gmres.cpp / gmres_perf_test.cpp are compiled by nvcc (their Kokkos lambdas become kernels through the shim), the
# reference's `Cuda` device name is bound to include/b200 (shim_b200/types_cuda.hpp), kernels_mkl.cpp keeps serving
# the host side (LoadMatrix, b = A x_true).  Output: _ref/gmres_perf_test_b200 ( --gpu = B200 backend ).
NVCC      = nvcc
NVFLAGS   = -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --extended-lambda -w -x cu -ccbin /usr/bin/g++ \
            -Xcompiler -fopenmp,-fPIC -I shim_b200 -I shim -I $(FARM) -I ../include
B200LIB   = ../icl-mixed-precision-gmres_b200/lib

b200: $(OUT)/gmres_perf_test_b200

$(OUT)/b200_gmres.o: $(FARM)/.linked shim/Kokkos_Core.hpp ../include/b200/types_b200.hpp
	$(NVCC) $(NVFLAGS) -c $(FARM)/gmres.cpp -o $@
$(OUT)/b200_main.o: $(FARM)/.linked shim/Kokkos_Core.hpp ../include/b200/types_b200.hpp
	$(NVCC) $(NVFLAGS) -include cstring -include sstream -c $(FARM)/gmres_perf_test.cpp -o $@
$(OUT)/b200_kernels.o: ../include/b200/kernels_b200.cpp ../include/b200/types_b200.hpp shim/Kokkos_Core.hpp $(FARM)/.linked
	$(NVCC) $(NVFLAGS) -c ../include/b200/kernels_b200.cpp -o $@
$(OUT)/b200_kokkos_shim.o: shim/kokkos_cuda_shim.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O2 -ccbin /usr/bin/g++ -Xcompiler -fPIC -c $< -o $@
$(OUT)/b200_kernels_mkl.o: $(FARM)/.linked shim/Kokkos_Core.hpp shim/mkl.h
	$(CXX) -O2 -march=x86-64-v3 -fopenmp -w -std=c++14 -fPIC -I shim_b200 -I shim -I $(FARM) -I ../include -c $(FARM)/kernels_mkl.cpp -o $@

$(OUT)/gmres_perf_test_b200: $(OUT)/b200_main.o $(OUT)/b200_gmres.o $(OUT)/b200_kernels.o $(OUT)/b200_kokkos_shim.o $(OUT)/b200_kernels_mkl.o $(OUT)/mkl_shim.o $(OUT)/mmio.o
	$(NVCC) -ccbin /usr/bin/g++ -o $@ $^ -L$(B200LIB) -lmpgmres_b200 -Xlinker -rpath,'$$ORIGIN/../../icl-mixed-precision-gmres_b200/lib' \
	    -L$(TORCHLIB) -ltorch_cpu -lc10 -Xlinker -rpath,$(TORCHLIB) -lgomp

.PHONY: b200
