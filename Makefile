# Builds libbluetangle_cuda.so (sm_100a) and the C oracle.  Used by __graft_entry__.build().
NVCC      ?= nvcc
PKG       := bluetangle.jl_b200
CSRC      := $(wildcard $(PKG)/csrc/*.cu)
OBJ       := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(CSRC))
LIB       := $(PKG)/lib/libbluetangle_cuda.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v

all: $(LIB) oracle

$(LIB): $(OBJ)
	@mkdir -p $(PKG)/lib
	$(NVCC) -shared -o $@ $(OBJ)

build/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh) include/bluetangle_cuda.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

# bt_tile.cu goes through nvcc's own steps with one extra pass over the PTX: the micro-op switch of the register programs
# becomes an indirect branch (tools/ptx_brx.py; BT_NO_BRX=1 builds it with plain nvcc)
ifndef BT_NO_BRX
build/bt_tile.o: $(PKG)/csrc/bt_tile.cu $(wildcard $(PKG)/csrc/*.cuh) include/bluetangle_cuda.h tools/nvcc_brx.py tools/ptx_brx.py
	@mkdir -p build
	python3 tools/nvcc_brx.py $(NVFLAGS) -c $< -o $@ 2> build/bt_tile.ptxas.log || (cat build/bt_tile.ptxas.log; exit 1)
endif

oracle: oracle/_build/libbt_oracle_c.so

oracle/_build/libbt_oracle_c.so: oracle/strided_cpu.c
	@mkdir -p oracle/_build
	gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all oracle clean
