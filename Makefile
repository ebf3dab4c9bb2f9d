# Builds the C-ABI library (CUDA kernels for sm_100a + host half) and the host tools, in-tree.
#   make            -> gappadder_b200/libgappadder_b200.so, build/ContigsMerger_b200, build/TERefiner_b200, build/microbench_int
#   make oracle     -> oracle/_build/liboverlap_oracle.so (+ oracle/_ref/ when /root/reference exists)
NVCC ?= nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall -Iinclude
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wextra -Iinclude
CSRC := gappadder_b200/csrc
LIB := gappadder_b200/libgappadder_b200.so

HOST := gappadder_b200/host
all: $(LIB) build/microbench_int build/ContigsMerger_b200 build/TERefiner_b200

build:
	mkdir -p build

build/gp_api.o: $(CSRC)/gp_api.cu $(wildcard $(CSRC)/*.cuh) include/gappadder_b200.h | build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/gp_api.ptxas.log || (cat build/gp_api.ptxas.log; false)

build/int_peak.o: $(CSRC)/int_peak.cu include/gappadder_b200.h | build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/gp_host.o: $(CSRC)/gp_host.cpp include/gappadder_b200.h | build
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): build/gp_api.o build/int_peak.o build/gp_host.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

build/ContigsMerger_b200: $(HOST)/contigs_merger_main.cpp $(HOST)/merger.cpp $(HOST)/merge_graph.cpp $(HOST)/fasta.cpp $(HOST)/server.cpp $(HOST)/dedup.cpp $(HOST)/dedup.hpp $(HOST)/device_gate.hpp $(HOST)/merger.hpp $(HOST)/merge_graph.hpp $(HOST)/fasta.hpp $(HOST)/server.hpp $(LIB)
	$(CXX) $(CXXFLAGS) -o $@ $(HOST)/contigs_merger_main.cpp $(HOST)/merger.cpp $(HOST)/merge_graph.cpp $(HOST)/fasta.cpp $(HOST)/server.cpp $(HOST)/dedup.cpp -Lgappadder_b200 -lgappadder_b200 -Wl,-rpath,'$$ORIGIN/../gappadder_b200' -lpthread

build/TERefiner_b200: $(HOST)/terefiner_main.cpp $(HOST)/local_alignment.cpp $(HOST)/local_alignment.hpp $(LIB)
	$(CXX) $(CXXFLAGS) -o $@ $(HOST)/terefiner_main.cpp $(HOST)/local_alignment.cpp -Lgappadder_b200 -lgappadder_b200 -Wl,-rpath,'$$ORIGIN/../gappadder_b200' -lpthread

build/microbench_int: tools/microbench_int.cu | build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

# diagnostic library with phase stamps in the certificate kernel (tools/lone_pair_bench.py --trace); never the product
trace: | build
	$(NVCC) $(NVFLAGS) -DGP_WF16C_TRACE -c $(CSRC)/gp_api.cu -o build/gp_api_trace.o
	$(NVCC) $(ARCH) -shared -o build/libgappadder_b200_trace.so build/gp_api_trace.o build/int_peak.o build/gp_host.o -cudart static

oracle:
	$(MAKE) -C oracle all
	./oracle/build_ref.sh

clean:
	rm -rf build $(LIB)
.PHONY: all oracle clean trace
