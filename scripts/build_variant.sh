#!/bin/bash
# build_variant.sh NAME [-DFLAG=..]...: compile an A/B variant of the library into build_variants/libgx_NAME.so
# (select it at run time with GALAX_B200_LIB=build_variants/libgx_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
python - "$name" "$@" <<'PY'
import sys
from pathlib import Path
sys.path.insert(0, ".")
from galax_b200 import _lib
print(_lib.build(defines=sys.argv[2:], out=Path("build_variants") / f"libgx_{sys.argv[1]}.so"))
PY
