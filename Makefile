# Convenience wrapper around __graft_entry__.build(): the CUDA library (sm_100a only), the CPU checkers, the host-compiled
# kernel source for CPU tests.  `make test` runs the CPU suite; GPU tests need a B200 (`python -m pytest tests -m gpu`).
NVCC ?= nvcc
LIB = plonk.c_b200/libplonk_b200.so
CSRC = $(wildcard plonk.c_b200/csrc/*.cu plonk.c_b200/csrc/*.cuh) $(wildcard include/*.h)

all: $(LIB) oracle hostcheck

$(LIB): $(CSRC)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include \
	    -o $@ plonk.c_b200/csrc/cabi.cu plonk.c_b200/csrc/dropin.cu

oracle: $(LIB)
	$(MAKE) -C oracle all

hostcheck:
	python -c "import __graft_entry__ as g; g.build_hostcheck()"

test: all
	python -m pytest tests -x -q -m "not gpu"

clean:
	rm -f $(LIB) tests/hostcheck/*.so
	$(MAKE) -C oracle clean

.PHONY: all oracle hostcheck test clean
