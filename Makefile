# Convenience targets; the driver's entry point is __graft_entry__.build().
PY ?= python

all: lib            ## librawhash_b200.so (nvcc, sm_100a) + rawhash2_b200 (g++)
lib:
	$(PY) -m rawhash_b200.build
checkers:           ## oracle/librh_oracle.so and, when /root/reference exists, oracle/_ref/* (reference, slow5lib, rawhash2)
	$(MAKE) -C oracle all
dropin: lib checkers ## oracle/_ref/rawhash2_gpu: the reference's own rawhash2 with the kt_for line replaced
	$(PY) integration/build_dropin.py
test:               ## CPU suite (the GPU suite: pytest tests -m gpu on a B200)
	$(PY) -m pytest tests -x -q -m "not gpu"
clean:
	rm -rf rawhash_b200/csrc/_obj rawhash_b200/librawhash_b200.so rawhash_b200/rawhash2_b200
	$(MAKE) -C oracle clean

.PHONY: all lib checkers dropin test clean
