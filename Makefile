# Builds the product library (C ABI of include/*.h), the stcsp command-line tool and the test oracle.
#   make            -> stcsp_solver_b200/libstcsp_b200.so, bin/stcsp, oracle/liboracle.so (+ oracle/_ref when the
#                      reference sources are present)
# sm_100a only: the kernels are written for B200.
NVCC ?= nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
INC := -Iinclude -Istcsp_solver_b200/csrc/host -Istcsp_solver_b200/csrc/gpu
NVFLAGS := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unused-parameter $(INC)
CXXFLAGS := -std=c++17 -O2 -fPIC -Wall -Wextra $(INC) -I/usr/local/cuda/include
B := build

HOST_SRC := $(wildcard stcsp_solver_b200/csrc/host/*.cpp)
GPU_CPP := stcsp_solver_b200/csrc/gpu/compile.cpp
GPU_CU := stcsp_solver_b200/csrc/gpu/kernels.cu stcsp_solver_b200/csrc/gpu/solver.cu stcsp_solver_b200/csrc/gpu/automaton.cu \
          stcsp_solver_b200/csrc/gpu/exchange.cu
OBJ := $(patsubst %.cpp,$(B)/%.o,$(notdir $(HOST_SRC) $(GPU_CPP))) $(patsubst %.cu,$(B)/%.o,$(notdir $(GPU_CU)))
HDR := $(wildcard include/*.h stcsp_solver_b200/csrc/host/*.h stcsp_solver_b200/csrc/gpu/*.h stcsp_solver_b200/csrc/gpu/*.cuh)
LIB := stcsp_solver_b200/libstcsp_b200.so

all: $(LIB) bin/stcsp oracle

$(B)/%.o: stcsp_solver_b200/csrc/host/%.cpp $(HDR) | $(B)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(B)/%.o: stcsp_solver_b200/csrc/gpu/%.cpp $(HDR) | $(B)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(B)/%.o: stcsp_solver_b200/csrc/gpu/%.cu $(HDR) | $(B)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(B):
	mkdir -p $(B) bin

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart shared -ldl

bin/stcsp: stcsp_solver_b200/csrc/cli/main.cpp $(LIB) $(HDR)
	$(CXX) $(CXXFLAGS) -o $@ $< -Lstcsp_solver_b200 -lstcsp_b200 -Wl,-rpath,'$$ORIGIN/../stcsp_solver_b200'

oracle: $(LIB)
	$(MAKE) -C oracle

clean:
	rm -rf $(B) $(LIB) bin/stcsp
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
