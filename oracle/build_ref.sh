#!/usr/bin/env bash
# ORACLE tooling: compile the reference's ROIAlign CPU source where it lies (never copied) into
# oracle/_ref/libref_roialign.so.  Build container only (needs /root/reference).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VETO_REFERENCE_ROOT:-/root/reference}"
[ -f "$REF/pysgg/csrc/cpu/ROIAlign_cpu.cpp" ] || { echo "reference not present, skipping oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
PY=python
TORCH_DIR=$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')
PY_INC=$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')
g++ -O2 -std=c++17 -shared -fPIC -w -include "$HERE/ref_compat.h" \
    -I"$REF/pysgg/csrc" -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PY_INC" \
    -D_GLIBCXX_USE_CXX11_ABI=$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))') \
    "$REF/pysgg/csrc/cpu/ROIAlign_cpu.cpp" "$HERE/ref_binding.cpp" \
    -L"$TORCH_DIR/lib" -Wl,-rpath,"$TORCH_DIR/lib" -ltorch -ltorch_cpu -lc10 \
    -o "$HERE/_ref/libref_roialign.so"
echo "built $HERE/_ref/libref_roialign.so"
