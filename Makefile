# Convenience targets; the build logic lives in tools/build_native.py.
.PHONY: all tests sass clean cpu-test
all:
	python tools/build_native.py
tests:
	python tools/build_native.py --tests && build/test_core
sass:
	python tools/build_native.py --sass
cpu-test: all
	python -m pytest tests -q -m "not gpu"
clean:
	python tools/build_native.py --clean
