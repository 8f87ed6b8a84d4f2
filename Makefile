# Top-level build: the product (libnw_cuda.so, sm_100a only), the reference-side binding (cuda.e, only where the
# reference tree is present) and the checker (oracle/).  nvcc cross-compiles without a GPU.
PKG    := fast-needleman-wunsch_b200
NVCC   ?= /usr/local/cuda/bin/nvcc
CXX    := /usr/bin/g++
REF    ?= /root/reference
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS:= $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $(EXTRA)
LIB    := $(PKG)/libnw_cuda.so
SRCS   := $(PKG)/csrc/nw_cuda.cu
HDRS   := $(wildcard $(PKG)/csrc/*.cuh) include/nw_cuda.h

.PHONY: all lib driver oracle clean check
all: lib oracle driver check

lib: $(LIB)
$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS)

# the same library with every global-memory index of the strip kernels asserted (NW_ASSERT in csrc/nw_kernels.cuh):
# the in-tree substitute for compute-sanitizer; tests/test_gpu_checked.py runs tools/sanity_small.py against it
check: build/libnw_check.so
build/libnw_check.so: $(SRCS) $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DNW_CHECK=1 -shared -o $@ $(SRCS)

oracle:
	$(MAKE) -C oracle all REF=$(REF)

# the reference's own driver + loader, unchanged, around our entry point (needs the reference tree; the built
# binary travels to the GPU box, where the tree is absent)
ifneq ($(wildcard $(REF)/src/common/driver.cpp),)
driver: $(PKG)/bin/cuda.e
$(PKG)/bin/cuda.e: $(PKG)/csrc/cuda.cpp $(LIB) oracle
	@mkdir -p $(PKG)/bin
	$(CXX) -Wall -std=c++11 -O3 -Wno-mismatched-new-delete -I $(REF)/src/common -I include -o $@ $< oracle/_ref/helper.o \
	    -L$(PKG) -lnw_cuda -Wl,-rpath,'$$ORIGIN/..'
else
driver:
	@echo "reference tree $(REF) not present: keeping prebuilt $(PKG)/bin/cuda.e as is"
endif

clean:
	rm -f $(LIB) $(PKG)/bin/cuda.e
	$(MAKE) -C oracle clean
