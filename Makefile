# Builds the product library (CUDA, sm_100a) and the CPU oracle (test infrastructure).
#   make            -> resvg_b200/libresvg_b200.so + oracle/liboracle.so
#   make lib        -> product only
#   make oracle     -> oracle only
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference (Rust) never contracts a*b+c; no -use_fast_math: IEEE div/sqrt.
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off \
             -Xptxas -v --expt-relaxed-constexpr -Wno-deprecated-gpu-targets
CSRC      := resvg_b200/csrc
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/*.cpp)
CU_OBJS   := $(CU_SRCS:.cu=.o)
CPP_OBJS  := $(CPP_SRCS:.cpp=.o)
HDRS      := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/resvg_b200.h

LIB       := resvg_b200/libresvg_b200.so
ORACLE    := oracle/liboracle.so
ORC_SRCS  := $(wildcard oracle/*.c)

all: lib oracle
lib: $(LIB)
oracle: $(ORACLE)

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(CSRC)/%.o: $(CSRC)/%.cpp $(HDRS)
	g++ -O2 -std=c++17 -fPIC -ffp-contract=off -I/usr/local/cuda/include -c $< -o $@

$(LIB): $(CU_OBJS) $(CPP_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread

# The oracle restates Rust f32/f64 arithmetic: no contraction, no fast-math.
$(ORACLE): $(ORC_SRCS) oracle/oracle.h
	$(CC) -O2 -std=c11 -fPIC -shared -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function \
	      -o $@ $(ORC_SRCS) -lm

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log $(LIB) $(ORACLE)

.PHONY: all lib oracle clean
