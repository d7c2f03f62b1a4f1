# Builds librakau_b200.so (sm_100a only) and the oracle. `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH      = -gencode arch=compute_100a,code=sm_100a
NVFLAGS   = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr
CSRC      = rakau_b200/csrc
OBJDIR    = build
LIB       = rakau_b200/lib/librakau_b200.so
OBJS      = $(OBJDIR)/sort.o $(OBJDIR)/build.o $(OBJDIR)/traverse.o $(OBJDIR)/leapfrog.o $(OBJDIR)/capi.o $(OBJDIR)/plummer.o

all: $(LIB) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/scan.cuh include/rakau_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(OBJDIR)/plummer.o: $(CSRC)/plummer.cpp include/rakau_b200.h
	@mkdir -p $(OBJDIR)
	g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -pthread -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p rakau_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# (oracle/_ref/libref_bridge.so links against the product library)
oracle: $(LIB)
	$(MAKE) -C oracle all

clean:
	rm -rf $(OBJDIR) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
