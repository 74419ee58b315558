# Builds the C-ABI library (hand-written sm_100a kernels) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` does the same.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := legged_gym_dev_b200/csrc
SRCS := $(wildcard $(CSRC)/*.cu)
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB  := legged_gym_dev_b200/libb200gym.so
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/b200gym.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)
.PHONY: all clean
