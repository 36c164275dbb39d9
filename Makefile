# Top-level build.
#   make            -> x264-dsp_b200/libx264dsp_b200.so   (the product: CUDA, sm_100a only)
#   make oracle     -> oracle/_build/libx264dsp_oracle.so (CPU checker, test infrastructure)
#   make ref        -> oracle/_ref/libx264ref.so          (unmodified reference, needs /root/reference)
#   make sass       -> x264-dsp_b200/_build/*.sass        (cuobjdump -sass of every kernel)
#   make glue       -> glue/_build/x264ref_gpu             (reference CLI + glue/*.c + the product, no Python)
#   make examples   -> examples/_build/lookahead_host      (plain C host program on the C ABI, gcc only)

NVCC     ?= /usr/local/cuda/bin/nvcc
PKG      := x264-dsp_b200
CSRC     := $(PKG)/csrc
OUT      := $(PKG)/_build
LIB      := $(PKG)/libx264dsp_b200.so

NVFLAGS  := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
            -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v -Iinclude
CU_SRCS  := $(wildcard $(CSRC)/*.cu)
CPP_SRCS := $(wildcard $(CSRC)/*.cpp)
OBJS     := $(CU_SRCS:$(CSRC)/%.cu=$(OUT)/%.o) $(CPP_SRCS:$(CSRC)/%.cpp=$(OUT)/%.o)
HDRS     := $(wildcard $(CSRC)/*.cuh) $(wildcard include/*.h)

all: $(LIB)

$(OUT)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OUT)/$*.ptxas.raw || (cat $(OUT)/$*.ptxas.raw; false)
	@grep -v "Compile time" $(OUT)/$*.ptxas.raw > $(OUT)/$*.ptxas.log; rm -f $(OUT)/$*.ptxas.raw

$(OUT)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OUT)
	g++ -O2 -fPIC -std=c++17 -Wall -Iinclude -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJS) -lpthread

oracle:
	$(MAKE) -C oracle

ref:
	$(MAKE) -C oracle ref

# the unmodified reference CLI linked with glue/*.c and $(LIB): glue/_build/x264ref_gpu (needs /root/reference)
glue: $(LIB)
	$(MAKE) -C glue

examples: examples/_build/lookahead_host

examples/_build/lookahead_host: examples/lookahead_host.c include/x264dsp_b200.h $(LIB)
	@mkdir -p examples/_build
	gcc -std=c99 -O2 -Wall -Werror -Iinclude $< -o $@ -L$(PKG) -l:libx264dsp_b200.so -Wl,-rpath,'$$ORIGIN/../../$(PKG)'

sass: $(LIB)
	/usr/local/cuda/bin/cuobjdump -sass $(LIB) > $(OUT)/libx264dsp_b200.sass

clean:
	rm -rf $(OUT) $(LIB) examples/_build
	$(MAKE) -C oracle clean

.PHONY: all oracle ref glue sass clean examples
