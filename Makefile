# Builds libpacmensl_b200.so (CUDA kernels + C ABI + host C++ mirror of the reference API) for sm_100a,
# and the CPU oracle.  `make` here cross-compiles without a GPU.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX_HOST := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(CXX_HOST) -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function,-Wno-deprecated-declarations
CXXFLAGS := -O2 -g -std=c++17 -fPIC -Wall -Wno-unused-function -fvisibility=hidden -Iinclude -Ipacmensl_b200/host -Ipacmensl_b200/fixtures

BUILD := build
LIB := pacmensl_b200/lib/libpacmensl_b200.so

CU_SRCS := $(wildcard pacmensl_b200/csrc/*.cu)
CU_OBJS := $(patsubst pacmensl_b200/csrc/%.cu,$(BUILD)/%.cu.o,$(CU_SRCS))
HOST_SRCS := $(wildcard pacmensl_b200/host/*.cpp)
HOST_OBJS := $(patsubst pacmensl_b200/host/%.cpp,$(BUILD)/%.host.o,$(HOST_SRCS))

all: $(LIB) oracle cpptests examples

$(BUILD)/%.cu.o: pacmensl_b200/csrc/%.cu pacmensl_b200/csrc/fsp_common.cuh include/fsp_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.host.o: pacmensl_b200/host/%.cpp $(wildcard pacmensl_b200/host/*.h) $(wildcard include/*.h) pacmensl_b200/fixtures/fsp_models.h
	@mkdir -p $(BUILD)
	$(CXX_HOST) $(CXXFLAGS) -c $< -o $@

# dense 62x62 expm on the host sits between two basis generations of KrylovFsp: optimise it harder
$(BUILD)/arma_shim.host.o: CXXFLAGS += -O3

$(LIB): $(CU_OBJS) $(HOST_OBJS)
	@mkdir -p pacmensl_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart_static -ldl -lpthread -lrt

oracle:
	$(MAKE) -s -C oracle

# C++ host-layer test programs (shaped after the reference's gtest programs); run by tests/test_gpu_host_cpp.py
CPP_TESTS := $(patsubst tests/cpp/%.cpp,$(BUILD)/tests/%,$(wildcard tests/cpp/test_*.cpp))
EXAMPLES := $(patsubst examples/%.cpp,$(BUILD)/examples/%,$(wildcard examples/*.cpp))
$(BUILD)/tests/%: tests/cpp/%.cpp tests/cpp/mini_gtest.h tests/cpp/pacmensl_test_env.h $(LIB)
	@mkdir -p $(BUILD)/tests
	$(CXX_HOST) -O1 -g -std=c++17 -Wall -Wno-unused-function -Iinclude -Ipacmensl_b200/host -Ipacmensl_b200/fixtures -Itests/cpp $< -o $@ -Lpacmensl_b200/lib -lpacmensl_b200 -Wl,-rpath,'$$ORIGIN/../../pacmensl_b200/lib'
$(BUILD)/examples/%: examples/%.cpp examples/example_common.h $(LIB)
	@mkdir -p $(BUILD)/examples
	$(CXX_HOST) -O2 -g -std=c++17 -Wall -Wno-unused-function -Iinclude -Ipacmensl_b200/host -Ipacmensl_b200/fixtures $< -o $@ -Lpacmensl_b200/lib -lpacmensl_b200 -Wl,-rpath,'$$ORIGIN/../../pacmensl_b200/lib'
cpptests: $(CPP_TESTS)
examples: $(EXAMPLES)

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c pacmensl_b200/csrc/fspmat.cu -o /dev/null

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info cpptests examples
