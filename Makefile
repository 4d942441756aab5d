# Builds the product library  russell_b200/lib/libsolver_b200.so  (sm_100a only) and the test-only oracle libs.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function
CXXFLAGS = -O2 -std=c++17 -fPIC -Wall -pthread
CSRC = russell_b200/csrc
LIB = russell_b200/lib/libsolver_b200.so
OBJ = build/solver_b200.o build/complex_b200.o build/symbolic.o build/ordering.o build/matching.o build/host_formats.o

all: $(LIB) oracle

$(LIB): $(OBJ)
	mkdir -p russell_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static -lpthread

build/solver_b200.o: $(CSRC)/solver_b200.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/solver_b200.h
	mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/ptxas_solver_b200.log || (cat build/ptxas_solver_b200.log; false)

build/complex_b200.o: $(CSRC)/complex_b200.cu $(wildcard $(CSRC)/*.hpp) include/solver_b200.h
	mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/%.o: $(CSRC)/%.cpp $(CSRC)/plan.hpp
	mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
