# Builds libclownresampler_b200.so (sm_100a only) in-tree: clownresampler_b200/lib/
NVCC ?= /usr/local/cuda/bin/nvcc
CC ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -Xcompiler -fPIC -std=c++17
CFLAGS := -O2 -fPIC -std=gnu99 -Wall -Wextra -Wno-unused-parameter

SRC := clownresampler_b200/csrc
OUT := clownresampler_b200/lib
OBJ := build/obj

LIB := $(OUT)/libclownresampler_b200.so

.PHONY: all clean oracle ref dropin voices-bench
all: $(LIB)

# the tiled kernel is instantiated one (kernel kind, channel range) per translation unit: `make -j` builds them in parallel
INST_KINDS := 0 1 6 8 10 12
INST_OBJ := $(foreach k,$(INST_KINDS),$(OBJ)/crb_inst_k$(k)_p0.o $(OBJ)/crb_inst_k$(k)_p1.o) $(foreach k,0 1,$(OBJ)/crb_inst_k$(k)_p2.o $(OBJ)/crb_inst_k$(k)_p3.o)

$(OBJ)/crb_device.o: $(SRC)/crb_device.cu $(SRC)/crb_kernels.cuh $(SRC)/crb_internal.h
	mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
define INST_RULE
$(OBJ)/crb_inst_k$(1)_p$(2).o: $(SRC)/crb_inst.cu $(SRC)/crb_kernels.cuh $(SRC)/crb_internal.h
	mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DCRB_INST_K=$(1) -DCRB_INST_PART=$(2) -c -o $$@ $$<
endef
$(foreach k,$(INST_KINDS),$(eval $(call INST_RULE,$(k),0)) $(eval $(call INST_RULE,$(k),1)))
$(foreach k,0 1,$(eval $(call INST_RULE,$(k),2)) $(eval $(call INST_RULE,$(k),3)))
$(OBJ)/crb_api.o: $(SRC)/crb_api.c $(SRC)/crb_internal.h include/clownresampler.h include/clownresampler_b200.h
	mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c -o $@ $<
$(OBJ)/crb_plan.o: $(SRC)/crb_plan.c $(SRC)/crb_internal.h
	mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c -o $@ $<
$(OBJ)/crb_voices.o: $(SRC)/crb_voices.c $(SRC)/crb_internal.h include/clownresampler.h include/clownresampler_b200.h
	mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c -o $@ $<

# only the public API (ClownResampler_* / ClownResamplerB200_*) is exported; the crb_* glue stays internal
$(OBJ)/exports.map: Makefile
	mkdir -p $(OBJ)
	printf '{ global: ClownResampler_*; ClownResamplerB200_*; local: *; };\n' > $@

$(LIB): $(OBJ)/crb_device.o $(INST_OBJ) $(OBJ)/crb_api.o $(OBJ)/crb_plan.o $(OBJ)/crb_voices.o $(OBJ)/exports.map
	mkdir -p $(OUT)
	$(NVCC) $(ARCH) -shared -o $@ $(filter %.o,$^) -Xlinker --version-script=$(OBJ)/exports.map -lpthread -lm

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

# The reference's own test programs, UNMODIFIED, compiled against include/clownresampler.h and linked
# with the CUDA library (build container only: the sources are read where they lie and only
# binaries are written, into oracle/_ref/).  They include "../clownresampler.h" and "dr_flac.h"
# relative to their own directory, so a scratch tree of symlinks stands in for that layout.
REFERENCE_DIR ?= /root/reference
dropin: $(LIB)
	mkdir -p build/dropin/tests oracle/_ref
	ln -sf $(abspath include/clownresampler.h) build/dropin/clownresampler.h
	ln -sf $(REFERENCE_DIR)/tests/dr_flac.h build/dropin/tests/dr_flac.h
	ln -sf $(REFERENCE_DIR)/tests/test-low-level.c build/dropin/tests/test-low-level.c
	ln -sf $(REFERENCE_DIR)/tests/test-high-level.c build/dropin/tests/test-high-level.c
	$(CC) -O2 -w -o oracle/_ref/dropin-test-low-level build/dropin/tests/test-low-level.c -L$(OUT) -lclownresampler_b200 -Wl,-rpath,'$$ORIGIN/../../$(OUT)' -lm
	$(CC) -O2 -w -o oracle/_ref/dropin-test-high-level build/dropin/tests/test-high-level.c -L$(OUT) -lclownresampler_b200 -Wl,-rpath,'$$ORIGIN/../../$(OUT)' -lm

# config-4 benchmark harness (many HighLevel voices): the same source against the reference and against us
voices-bench: $(LIB)
	mkdir -p oracle/_ref
	$(CC) -O2 -w -Iinclude -DWITH_BATCH -o $(OUT)/bench-highlevel tools/bench_highlevel.c -L$(OUT) -lclownresampler_b200 -Wl,-rpath,'$$ORIGIN' -lm
	if [ -d $(REFERENCE_DIR) ]; then $(CC) -O2 -w -DUSE_REFERENCE -I$(REFERENCE_DIR) -o oracle/_ref/bench-highlevel-ref tools/bench_highlevel.c -lm; fi

clean:
	rm -rf build $(OUT)
