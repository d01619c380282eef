#!/bin/bash
# CPU suite with coverage of the Python layer (reference counterpart: run_test.sh)
python tools/build_native.py --tests && build/test_core && \
python -m pytest tests -q -m "not gpu" "$@"
