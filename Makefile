# Top-level build: the product library (sm_100a only), the oracle (test infrastructure) and, when the
# reference tree is present (authoring container), the reference-hosted binaries.
NVCC   ?= nvcc
ARCH    = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
CSRC    = paml_b200/csrc
LIB     = paml_b200/libpaml_b200.so

REF    ?= /root/reference/src
HOST    = paml_b200/host
BIN     = paml_b200/bin
CC     ?= gcc

all: $(LIB) oracle host

# The drop-in programs: the unmodified reference driver (compiled from REF, nothing copied) + our lfun/lfundG.
# Built only where the reference tree exists; the binaries travel to the GPU box like the .so files do.
ifneq ($(wildcard $(REF)/codeml.c),)
host: $(BIN)/codeml_b200 $(BIN)/baseml_b200 $(BIN)/mcmctree_b200
else
host:
	@echo "host: reference sources not present at $(REF); using prebuilt $(BIN) if any"
endif

$(BIN)/tools.o: $(REF)/tools.c
	@mkdir -p $(BIN)
	$(CC) -O3 -w -c -o $@ $<
$(BIN)/codeml_b200: $(HOST)/codeml_b200.c $(HOST)/wrap_b200.inc $(HOST)/plfun_b200.inc $(HOST)/pb200_gather.inc include/paml_b200.h $(BIN)/tools.o $(LIB)
	$(CC) -O3 -w -I$(REF) -I$(HOST) -Iinclude -o $@ $(HOST)/codeml_b200.c $(BIN)/tools.o -Lpaml_b200 -lpaml_b200 -Wl,-rpath,'$$ORIGIN/..' -lm
$(BIN)/baseml_b200: $(HOST)/baseml_b200.c $(HOST)/wrap_b200.inc $(HOST)/plfun_b200.inc $(HOST)/pb200_gather.inc include/paml_b200.h $(BIN)/tools.o $(LIB)
	$(CC) -O3 -w -I$(REF) -I$(HOST) -Iinclude -o $@ $(HOST)/baseml_b200.c $(BIN)/tools.o -Lpaml_b200 -lpaml_b200 -Wl,-rpath,'$$ORIGIN/..' -lm
# mcmctree with its exact sequence likelihood (usedata = 1) through the same C ABI (host/mcmctree_b200.c)
$(BIN)/mcmctree_b200: $(HOST)/mcmctree_b200.c include/paml_b200.h $(BIN)/tools.o $(LIB)
	$(CC) -O3 -w -I$(REF) -I$(HOST) -Iinclude -o $@ $(HOST)/mcmctree_b200.c $(BIN)/tools.o -Lpaml_b200 -lpaml_b200 -Wl,-rpath,'$$ORIGIN/..' -lm

$(LIB): $(CSRC)/api.cu $(CSRC)/kernels.cuh $(CSRC)/patterns.cpp $(CSRC)/plan.cpp $(CSRC)/plan.h $(CSRC)/prune4.cuh $(CSRC)/prune4v2.cuh $(CSRC)/prune_tma.cuh $(CSRC)/prune_ps.cuh $(CSRC)/cond_dmma.cuh $(CSRC)/eigen.cuh $(CSRC)/cond.cuh include/paml_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/api.cu $(CSRC)/plan.cpp $(CSRC)/patterns.cpp -ldl 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "error|warning" build_ptxas.log || true

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB) build_ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle clean host
