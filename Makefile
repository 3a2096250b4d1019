# Convenience targets; the build itself lives in cosma_b200/build.py (nvcc for sm_100a + g++ for the C++ host layer) and oracle/Makefile.
PY ?= python

.PHONY: all build test test-gpu bench miniapps clean
all: build

build:
	$(PY) __graft_entry__.py

test: build
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -q -m gpu

bench: build
	$(PY) bench.py

# C++ programs against libcosma.so (also built on demand by tests/test_z_cpp_api.py)
miniapps: build
	mkdir -p tests/cpp/bin
	g++ -O2 -std=c++17 -I include miniapp/cosma_miniapp.cpp -o tests/cpp/bin/cosma_miniapp -L cosma_b200/lib -lcosma -lcosma_b200 -Wl,-rpath,$(CURDIR)/cosma_b200/lib
	g++ -O2 -std=c++17 -I include miniapp/pxgemm_miniapp.cpp -o tests/cpp/bin/pxgemm_miniapp -L cosma_b200/lib -lcosma_pxgemm_cpp -lcosma_blacs_lite -lcosma -lcosma_b200 -Wl,-rpath,$(CURDIR)/cosma_b200/lib
	g++ -O2 -std=c++17 -I include miniapp/pxgemr2d_miniapp.cpp -o tests/cpp/bin/pxgemr2d_miniapp -L cosma_b200/lib -lcosma_pxgemm_cpp -lcosma_blacs_lite -lcosma -lcosma_b200 -Wl,-rpath,$(CURDIR)/cosma_b200/lib
	g++ -O2 -std=c++17 -I include miniapp/pxtran_miniapp.cpp -o tests/cpp/bin/pxtran_miniapp -L cosma_b200/lib -lcosma_pxgemm_cpp -lcosma_blacs_lite -lcosma -lcosma_b200 -Wl,-rpath,$(CURDIR)/cosma_b200/lib

clean:
	rm -rf cosma_b200/build cosma_b200/lib tests/cpp/bin oracle/_ref oracle/liboracle.so
