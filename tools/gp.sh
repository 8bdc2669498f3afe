#!/bin/bash
# rebuild the native artefacts (so that the snapshot ships a current .so), then run a command on the GPU box
# usage: tools/gp.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build_cuda(); g.build_oracle()" 2>&1 | grep -E "error|Error" && exit 1
python -c "
import __graft_entry__ as g, ctypes
mm = g.load_package()
from magic_mirror_b200 import _lib
h = ctypes.CDLL(mm.LIB_PATH)
missing = [n for n in _lib.SIGNATURES if not hasattr(h, n)]
assert not missing, missing
"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
