# Builds the C-ABI library (CUDA kernels for sm_100a + host table phase) in-tree.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-ffp-contract=off --fmad=false
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -ffp-contract=off
SRC       := contrack_b200/csrc
OUT       := contrack_b200/lib
LIB       := $(OUT)/libcontrack_b200.so

SYNTH     := bench_support/libct_synth.so

all: $(LIB) $(SYNTH)

# benchmark tooling only (synthetic input generator); not linked into the product library
$(SYNTH): bench_support/ct_synth.cu
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -o $@ $< -lcudart_static -lpthread -ldl -lrt

$(OUT)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h) include/contrack_b200.h
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OUT)/%.o: $(SRC)/%.cpp $(wildcard $(SRC)/*.h)
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(OUT)/ct_kernels.o $(OUT)/ct_plane.o $(OUT)/ct_global.o $(OUT)/ct_fast.o $(OUT)/ct_comm.o $(OUT)/ct_dist.o $(OUT)/ct_lifecycle.o $(OUT)/ct_shard.o $(OUT)/ct_extras.o $(OUT)/ct_anom.o $(OUT)/ct_api.o $(OUT)/ct_host.o $(OUT)/ct_tables.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart_static -lpthread -ldl -lrt

clean:
	rm -rf $(OUT) $(SYNTH)

.PHONY: all clean
