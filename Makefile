# Builds libseqkit_b200.so (sm_100a only) and the host `fasta` binary in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -diag-suppress 186
CSRC := seqkit_b200/csrc
LIB := seqkit_b200/libseqkit_b200.so
OBJ := $(CSRC)/sk_kernels.o $(CSRC)/sk_warp.o $(CSRC)/sk_compact.o $(CSRC)/sk_lineops.o $(CSRC)/sk_api.o $(CSRC)/sk_synth.o

all: $(LIB) seqkit_b200/fasta oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/sk_internal.h $(CSRC)/sk_device.cuh $(CSRC)/sk_record.cuh include/seqkit_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl

seqkit_b200/fasta: seqkit_b200/host/fasta_main.cpp $(LIB) include/seqkit_b200.h
	g++ -O2 -std=c++17 -Wall -Iinclude -o $@ $< -Lseqkit_b200 -lseqkit_b200 -Wl,-rpath,'$$ORIGIN' -lpthread -lz

# diagnostic build with per-phase cycle counters (tools/phase_timing.py)
seqkit_b200/libseqkit_b200_timing.so: $(CSRC)/sk_warp.cu $(CSRC)/sk_record.cuh $(CSRC)/sk_device.cuh $(CSRC)/sk_kernels.cu $(CSRC)/sk_compact.cu $(CSRC)/sk_lineops.cu $(CSRC)/sk_api.cu $(CSRC)/sk_synth.cu $(CSRC)/sk_internal.h include/seqkit_b200.h
	$(NVCC) $(NVFLAGS) -DSK_PHASE_TIMING -shared -o $@ $(CSRC)/sk_kernels.cu $(CSRC)/sk_warp.cu $(CSRC)/sk_compact.cu $(CSRC)/sk_lineops.cu $(CSRC)/sk_api.cu $(CSRC)/sk_synth.cu -ldl

oracle:
	$(MAKE) -s -C oracle all

clean:
	rm -f $(OBJ) $(LIB) seqkit_b200/fasta
	$(MAKE) -s -C oracle clean
.PHONY: all oracle clean

# A/B variants of the library built with extra -D switches (tools/runvar.sh; SK_LIB selects the .so):
#   make variant NAME=norun DEFS="-DSKW_EMIT_RUNS=0"
variant:
	mkdir -p seqkit_b200/variants
	$(NVCC) $(NVFLAGS) $(DEFS) -shared -o seqkit_b200/variants/$(NAME).so $(CSRC)/sk_kernels.cu $(CSRC)/sk_warp.cu $(CSRC)/sk_compact.cu $(CSRC)/sk_lineops.cu $(CSRC)/sk_api.cu $(CSRC)/sk_synth.cu -ldl
.PHONY: variant
