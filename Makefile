# Builds zk_b200/libzk_b200.so (C ABI in include/zk_b200.h) for sm_100a, in tree.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
SRC := zk_b200/csrc
OBJDIR := build
OBJS := $(OBJDIR)/api.o $(OBJDIR)/api_sumcheck.o $(OBJDIR)/api_ntt.o $(OBJDIR)/kernels_sumcheck.o $(OBJDIR)/kernels_sop.o $(OBJDIR)/kernels_mle.o $(OBJDIR)/kernels_ntt.o $(OBJDIR)/kernels_ntt_sharded.o $(OBJDIR)/microbench.o \
        $(OBJDIR)/keccak_avx512.o
HDRS := $(SRC)/field.cuh $(SRC)/field_f64.cuh $(SRC)/fold_imma.cuh $(SRC)/reduce.cuh $(SRC)/accw.cuh $(SRC)/sop_kernel.cuh $(SRC)/ntt_sharded_kernels.cuh $(SRC)/kernels.h $(SRC)/api_internal.h $(SRC)/host_field.hpp $(SRC)/keccak.hpp include/zk_b200.h

all: zk_b200/libzk_b200.so

$(OBJDIR)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

# host-only: the AVX-512 absorb loop of the transcript (selected at run time by CPUID)
$(OBJDIR)/keccak_avx512.o: $(SRC)/keccak_avx512.cpp
	@mkdir -p $(OBJDIR)
	$(CXX) -std=c++17 -O3 -mavx512f -mavx512vl -fPIC -fvisibility=hidden -c $< -o $@

zk_b200/libzk_b200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker --version-script=$(SRC)/exports.map -lcudart_static -ldl -lpthread -lrt

clean:
	rm -rf $(OBJDIR) zk_b200/libzk_b200.so

.PHONY: all clean
