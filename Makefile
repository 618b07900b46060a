# Build of the B200-native CRP-SpMM library (sm_100a only).
#
#   make            libminimpi.so, minimpirun, libcrpspmm.so           -> crp-spmm_b200/{lib,bin}
#   make drivers    the reference's example drivers, compiled UNCHANGED from $(REF)/examples
#                   against include/ + libcrpspmm.so (only where $(REF) exists)
#   make oracle     oracle/liboracle.so and, where $(REF) exists, oracle/_ref/*
#   make clean
REF     ?= /root/reference
ROOT    := $(abspath $(dir $(lastword $(MAKEFILE_LIST))))
PKG     := $(ROOT)/crp-spmm_b200
SRC     := $(PKG)/csrc
LIBDIR  := $(PKG)/lib
BINDIR  := $(PKG)/bin
OBJDIR  := $(ROOT)/build/obj
MINIMPI := $(PKG)/minimpi

CC      := $(or $(CRP_CC),gcc)
NVCC    := $(or $(CRP_NVCC),nvcc)
CFLAGS  := -O2 -g -std=gnu11 -fPIC -fopenmp -Wall -Wno-unused-function -I$(ROOT)/include -I$(MINIMPI) -I$(SRC)/host -I/usr/local/cuda/include
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function \
           -I$(ROOT)/include -I$(SRC)/cuda $(CRP_NVCC_EXTRA)

HOST_SRC := utils.c spmat_part.c dev_type.c crp_common.c rowpara_spmm.c para2d_spmm.c mat_redist.c crpspmm.c
COMM_SRC := crp_nccl.c
CUDA_SRC := $(notdir $(wildcard $(SRC)/cuda/*.cu))
OBJS := $(addprefix $(OBJDIR)/host_,$(HOST_SRC:.c=.o)) $(addprefix $(OBJDIR)/comm_,$(COMM_SRC:.c=.o)) $(addprefix $(OBJDIR)/cuda_,$(CUDA_SRC:.cu=.o))

.PHONY: all lib drivers oracle clean
all: lib

lib: $(LIBDIR)/libminimpi.so $(BINDIR)/minimpirun $(LIBDIR)/libcrpspmm.so $(LIBDIR)/libcrpingest.so

$(LIBDIR)/libcrpingest.so: $(SRC)/ingest/mmio_fast.c $(SRC)/ingest/graph_part.c $(SRC)/ingest/metis.h | $(LIBDIR)
	$(CC) -O3 -g -std=gnu11 -fPIC -fopenmp -Wall -I$(SRC)/ingest -shared -o $@ $(SRC)/ingest/mmio_fast.c $(SRC)/ingest/graph_part.c

$(OBJDIR) $(LIBDIR) $(BINDIR):
	mkdir -p $@

$(LIBDIR)/libminimpi.so: $(MINIMPI)/minimpi.c $(MINIMPI)/mpi.h | $(LIBDIR)
	$(CC) -O2 -g -fPIC -shared -o $@ $<

$(BINDIR)/minimpirun: $(MINIMPI)/minimpirun.c | $(BINDIR)
	$(CC) -O2 -o $@ $<

$(OBJDIR)/host_%.o: $(SRC)/host/%.c $(wildcard $(ROOT)/include/*.h) $(SRC)/host/crp_internal.h | $(OBJDIR)
	$(CC) $(CFLAGS) -c $< -o $@

$(OBJDIR)/comm_%.o: $(SRC)/comm/%.c $(wildcard $(ROOT)/include/*.h) $(SRC)/host/crp_internal.h | $(OBJDIR)
	$(CC) $(CFLAGS) -c $< -o $@

$(OBJDIR)/cuda_%.o: $(SRC)/cuda/%.cu $(wildcard $(SRC)/cuda/*.cuh) $(ROOT)/include/crp_cuda.h | $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libcrpspmm.so: $(OBJS) $(LIBDIR)/libminimpi.so
	$(NVCC) -shared -o $@ $(OBJS) -L$(LIBDIR) -lminimpi -ldl -lgomp -Xlinker -rpath='$$ORIGIN'

# ---- the reference's own drivers against this library (drop-in check) ----
DRV_INC  := -I$(ROOT)/include -I$(MINIMPI) -I$(SRC)/ingest -I$(ROOT)/oracle/stubs -I$(REF)/examples
# the drivers' matrix reader (mm_read_sparse_RPI + coo2csr of examples/mmio_utils.c) is replaced by the fast ingest library
# (same signatures and results: csrc/ingest/mmio_fast.c, tests/test_ingest.py) and METIS, which the reference links but does not
# vendor, by the native partitioner behind METIS's two entry points (csrc/ingest/graph_part.c); everything else is the reference's source
DRV_HELP := mmio.c test_utils.c metis_mat_part.c
DRV_OBJS := $(addprefix $(OBJDIR)/drv_,$(DRV_HELP:.c=.o)) $(OBJDIR)/drv_mkl_standin.o
DRV_CFLAGS := -O3 -march=x86-64-v3 -fopenmp -std=gnu11 -g -DUSE_MKL -Wno-unused-result

ifneq ($(wildcard $(REF)/examples/test_para2d_spmm.c),)
drivers: lib $(BINDIR)/test_para2d_spmm.exe $(BINDIR)/test_rp_spmm.exe $(BINDIR)/test_spmm_2dpg.exe $(BINDIR)/test_crpspmm.exe
else
drivers:
	@echo "drivers: $(REF) not present - keeping prebuilt drivers (if any)"
endif

$(OBJDIR)/drv_%.o: $(REF)/examples/%.c | $(OBJDIR)
	$(CC) $(DRV_CFLAGS) $(DRV_INC) -c $< -o $@
# the deprecated composite-engine driver, also unchanged (deprecated/examples/test_crpspmm.c) against include/crpspmm.h
$(OBJDIR)/drv_test_crpspmm.o: $(REF)/deprecated/examples/test_crpspmm.c | $(OBJDIR)
	$(CC) $(DRV_CFLAGS) $(DRV_INC) -c $< -o $@
$(OBJDIR)/drv_mkl_standin.o: $(ROOT)/oracle/stubs/mkl_standin.c | $(OBJDIR)
	$(CC) $(DRV_CFLAGS) -ffp-contract=off $(DRV_INC) -c $< -o $@

$(BINDIR)/%.exe: $(OBJDIR)/drv_%.o $(DRV_OBJS) $(LIBDIR)/libcrpspmm.so $(LIBDIR)/libcrpingest.so | $(BINDIR)
	$(CC) -fopenmp -o $@ $< $(DRV_OBJS) -L$(LIBDIR) -lcrpspmm -lcrpingest -lminimpi -lm -Wl,-rpath,'$$ORIGIN/../lib'

oracle:
	$(MAKE) -C $(ROOT)/oracle all

clean:
	rm -rf $(ROOT)/build $(LIBDIR) $(BINDIR)
.SECONDARY:
