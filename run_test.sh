#!/bin/bash
# CPU suite with line coverage of the Python layer (reference counterpart: run_test.sh, which
# runs `pytest --cov=infinistore`).  pytest-cov is used when it is installed; the offline image
# has neither pytest-cov nor coverage, so the in-repo sys.monitoring plugin stands in.
set -e
python tools/build_native.py --tests && build/test_core
if python -c "import pytest_cov" 2>/dev/null; then
    python -m pytest tests -q -m "not gpu" --cov=infinistore_b200 --cov-report=term "$@"
else
    python -m pytest tests -q -m "not gpu" -p tools.pycov --pycov "$@"
fi
