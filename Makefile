# Builds tools21cm_b200/lib/libt21ps.so (sm_100a) -- also what __graft_entry__.build() runs.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -fopenmp -Xcompiler -mavx2 --expt-relaxed-constexpr $(EXTRA_NVFLAGS)
SRC_DIR   := tools21cm_b200/csrc
BUILD_DIR ?= build/t21ps
LIB       ?= tools21cm_b200/lib/libt21ps.so
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(BUILD_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) $(wildcard $(SRC_DIR)/*.h) include/t21ps.h

all: $(LIB)

$(BUILD_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(BUILD_DIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static -Xcompiler -fopenmp

emu:
	g++ -O2 -std=c++17 -shared -fPIC -o tests/emu/libt21emu.so tests/emu/emu.cpp

clean:
	rm -rf $(BUILD_DIR) $(LIB)

.PHONY: all clean emu
