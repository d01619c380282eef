#!/bin/bash
# Builds the UNMODIFIED reference (bd-iaas-us/infiniStore, /root/reference) into
# baseline/_ref for `bench.py --impl reference`.
#
# The offline image lacks the reference's system dependencies (libuv-dev, flatbuffers,
# boost, rdma-core: src/Makefile:6).  What stands in for them - none of it touches the
# reference's own sources, which are compiled byte for byte with the reference's own Makefile:
#   libuv        the REAL libuv 1.48 exported by uvloop's extension module (the very loop object
#                the reference's server receives from uvloop); refshim/include/uv.h only
#                declares the API subset with the library's own struct sizes
#   spdlog/fmt   real headers vendored by flashinfer (spdlog with its bundled fmt)
#   flatbuffers  refshim/include/flatbuffers/flatbuffers.h: clean-room subset of the C++ API
#                that flatc-generated code uses, standard wire format
#   boost        refshim/include/boost/*: intrusive_ptr, lockfree::spsc_queue, stacktrace subsets
#   libibverbs   refshim/noverbs.c: a NULL provider - one pseudo device, PD/MR bookkeeping only,
#                no queue pairs.  The box has no RDMA hardware, so the reference's RDMA path
#                cannot run in any case; its LOCAL_GPU path (TCP + CUDA IPC + per-block
#                cudaMemcpyAsync, src/infinistore.cpp:570-804) uses no verbs and runs as is.
#   ibv_devinfo  refshim/bin/ibv_devinfo: prints "No IB devices found", exits 1 (what the real
#                tool does without hardware); check_supported() requires the command to exist.
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SHIM=$ROOT/baseline/refshim
REF=${REFERENCE_SRC:-/root/reference}
WORK=$(mktemp -d /tmp/ref_copy.XXXXXX)
PY=${PYTHON:-python}
SP=$($PY -c "import sysconfig; print(sysconfig.get_paths()['purelib'])")
UVSO=$($PY -c "import uvloop, glob, os; print(glob.glob(os.path.join(os.path.dirname(uvloop.__file__), 'loop*.so'))[0])")
SPDLOG=$SP/flashinfer/data/spdlog/include
[ -f "$SPDLOG/spdlog/spdlog.h" ] || { echo "spdlog headers not found under $SPDLOG"; exit 2; }
cp -r "$REF"/. "$WORK"/
cd "$WORK"
rm -f src/*.o src/*.so infinistore/*.so
git init -q . && git add -A >/dev/null && git -c user.email=ref@local -c user.name=ref commit -qm ref && git tag 0.0.0
gcc -O2 -fPIC -I"$SHIM/include" -c "$SHIM/noverbs.c" -o "$WORK/noverbs.o"
make -C src -j8 PYTHON="$PY" \
    INCLUDES="-I$SHIM/include -I$SPDLOG -I/usr/local/cuda/include" \
    LDFLAGS="-L/usr/local/cuda/lib64 -rdynamic -L$(dirname "$UVSO") -Wl,-rpath,$(dirname "$UVSO")" \
    LIBS="$WORK/noverbs.o -lcudart -l:$(basename "$UVSO") -ldl" \
    PYTHON_EXTENSION_SUFFIX="$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")" \
    PYBIND11_INCLUDES="$($PY -m pybind11 --includes)"
ls -la infinistore/_infinistore*.so
rm -rf "$ROOT/baseline/_ref"
$PY -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$WORK"
ls "$ROOT/baseline/_ref/infinistore"
cd /tmp && PYTHONPATH="$ROOT/baseline/_ref" $PY -c "import infinistore._infinistore as m; print('reference module imports:', m.__file__)"
rm -rf "$WORK"
