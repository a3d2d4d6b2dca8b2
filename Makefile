# Convenience targets (the driver uses __graft_entry__.build()/smoke(), pytest and bench.py directly).
PY ?= python

build:            ## nvcc -> hybrid-rag-colbertv2_b200/libhrc.so (sm_100a), gcc -> oracle/liboracle.so
	$(PY) -c "import __graft_entry__ as g; g.build()"
test-cpu:         ## oracle vs golden fixtures, host logic, ABI symbols (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"
test-gpu:         ## parity through the C ABI on a B200
	$(PY) -m pytest tests -q -m gpu
smoke:
	$(PY) -c "import __graft_entry__ as g; g.smoke()"
bench:            ## the BASELINE.json metric line (C2) with the `secondary` records for C1/C3/C4/sustained
	$(PY) bench.py
golden:           ## regenerate tests/golden from the unmodified reference (needs /root/reference)
	$(PY) tests/golden/make_golden.py
profile:          ## ncu launch list + full captures -> gpurun_out/, then copy into profiles/
	bash scripts/gpu_profile_r02.sh && $(PY) scripts/refresh_profiles.py r02
clean:
	$(MAKE) -C hybrid-rag-colbertv2_b200/csrc clean; rm -f oracle/liboracle.so
.PHONY: build test-cpu test-gpu smoke bench golden profile clean
