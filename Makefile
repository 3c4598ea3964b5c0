# Top-level build: the product library (sm_100a only), the oracle (test infrastructure) and, when the
# reference tree is present (authoring container), the reference-hosted binaries.
NVCC   ?= nvcc
ARCH    = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
CSRC    = paml_b200/csrc
LIB     = paml_b200/libpaml_b200.so

all: $(LIB) oracle

$(LIB): $(CSRC)/api.cu $(CSRC)/kernels.cuh $(CSRC)/plan.cpp $(CSRC)/plan.h include/paml_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/api.cu $(CSRC)/plan.cpp -ldl 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "error|warning" build_ptxas.log || true

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB) build_ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
