# Build of the product library (nvcc, sm_100a only) and of the test-side pieces.
#   make            -> iamr_b200/libiamrx.so        (CUDA, the only product path)
#   make oracle     -> oracle/_build/liboracle.so   (CPU restatement, test infrastructure)
#   make emul       -> tests/emul/_build/libiamrx_emul.so (host emulation harness, tests only)
NVCC      ?= /usr/local/cuda/bin/nvcc
# the environment exports CXX=/opt/gcc/bin/g++, which ships without libgomp; use the distro g++
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
SRCS      := $(wildcard iamr_b200/csrc/*.cu)
HDRS      := $(wildcard iamr_b200/csrc/*.h) include/iamrx.h
OBJS      := $(patsubst iamr_b200/csrc/%.cu,build/%.o,$(SRCS))
LIB       := iamr_b200/libiamrx.so

all: $(LIB)

build/%.o: iamr_b200/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -Xlinker -Bsymbolic -o $@ $(OBJS) -lcudart -ldl

oracle:
	$(MAKE) -C oracle

EMUL_OBJS := $(patsubst iamr_b200/csrc/%.cu,tests/emul/_build/%.o,$(SRCS))
emul: tests/emul/_build/libiamrx_emul.so
tests/emul/_build/%.o: iamr_b200/csrc/%.cu $(HDRS) tests/emul/cuda_emul.h
	@mkdir -p tests/emul/_build
	$(CXX) -x c++ -std=c++17 -O2 -fopenmp -fPIC -ftls-model=initial-exec -DIX_EMUL -Itests/emul -Wall -Wno-unused-function -Wno-unknown-pragmas -c $< -o $@
tests/emul/_build/libiamrx_emul.so: $(EMUL_OBJS) tests/emul/cuda_emul.cpp
	$(CXX) -std=c++17 -O2 -fopenmp -fPIC -shared -Wl,-Bsymbolic -DIX_EMUL -Itests/emul -o $@ $(EMUL_OBJS) tests/emul/cuda_emul.cpp -ldl

clean:
	rm -rf build $(LIB) tests/emul/_build oracle/_build

.PHONY: all oracle emul clean
