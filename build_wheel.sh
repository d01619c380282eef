#!/bin/bash
# Build a wheel that contains the prebuilt sm_100a native module (reference counterpart:
# build_manylinux_wheels.sh, which builds inside a manylinux container; here the module links
# the CUDA runtime statically and needs no NIC libraries, so a plain wheel is portable across
# hosts with the same CPython ABI).
set -euo pipefail
cd "$(dirname "$0")"
python tools/build_native.py
python setup.py -q bdist_wheel --dist-dir dist
ls -l dist/*.whl
