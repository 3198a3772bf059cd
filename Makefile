# Top-level build.  `make` builds everything that can be built on this box:
#   lib     photometric-bundle-adjustment_b200/libpba_b200.so  (nvcc, sm_100a; the product)
#   synth   photometric-bundle-adjustment_b200/libpba_synth.so (host-only scene generator)
#   oracle  oracle/libpba_oracle.so                             (CPU restatement; test infra)
#   ref     oracle/_ref/libpba_ref.so                           (unmodified reference + vendored
#                                                                Ceres; only when /root/reference exists)
PKG     := photometric-bundle-adjustment_b200
CSRC    := $(PKG)/csrc
NVCC    ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= g++
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-O3,-fvisibility=hidden,-fopenmp \
           -Iinclude -I$(CSRC) --expt-relaxed-constexpr -Xptxas -v $(EXTRA)
CU_SRCS := $(wildcard $(CSRC)/*.cu)
CU_HDRS := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/pba.h

all: lib synth oracle ref

# one object per translation unit (no relocatable device code is needed), so `make -j` compiles them in
# parallel; the per-file ptxas reports are concatenated into $(PKG)/ptxas.log
CU_OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS))
lib: $(PKG)/libpba_b200.so
build/%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
$(PKG)/libpba_b200.so: $(CU_OBJS)
	$(NVCC) $(ARCH) -shared $(CU_OBJS) -o $@ -lcudart -ldl -lgomp
	@cat build/*.ptxas.log > $(PKG)/ptxas.log

synth: $(PKG)/libpba_synth.so
$(PKG)/libpba_synth.so: $(CSRC)/synth.cpp $(CSRC)/pba_math.h $(CSRC)/synth_scene.h include/pba.h include/pba_synth.h
	$(HOSTCXX) -O2 -std=c++17 -fPIC -fopenmp -shared -Iinclude -I$(CSRC) $< -o $@

oracle: oracle/libpba_oracle.so
oracle/libpba_oracle.so: oracle/pba_oracle.cpp include/pba.h
	$(HOSTCXX) -O2 -std=c++17 -fPIC -fopenmp -fvisibility=hidden -shared -Iinclude $< -o $@

ref:
	@if [ -d /root/reference ]; then $(MAKE) -C oracle/ref both; else echo "no /root/reference: using prebuilt oracle/_ref if present"; fi

clean:
	rm -rf build
	rm -f $(PKG)/libpba_b200.so $(PKG)/libpba_synth.so oracle/libpba_oracle.so
.PHONY: all lib synth oracle ref clean
