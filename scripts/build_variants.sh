#!/bin/bash
# build the main library, then several variants in parallel:  scripts/build_variants.sh "name1 -DX=1" "name2 -DY=2 -DZ=3" ...
set -e
cd "$(dirname "$0")/.."
python -c "from deep_reconstruction_with_epipolar_lines_mvster_b200 import _build as B; B.build_library()"
rm -rf variants
for v in "$@"; do set -- $v; n=$1; shift; MVSTER_VARIANT_NO_MAIN=1 python scripts/build_variant.py $n -DMVSTER_FAST_BUILD "$@" -Xptxas -v > /tmp/var_$n.log 2>&1 & done
wait
ls variants/*.so
