#!/bin/bash
# usage: tools/ncu_by_function.sh <report.ncu-rep> <kernel tag, e.g. Li10ELi10ELi5ELi32> [launch index]
# Attributes ncu's per-instruction samples to device functions / source lines using nvdisasm line info
# of the in-tree library (must be the same build the report was captured with).
set -e
REP=$(realpath $1); TAG=$2; IDX=${3:-0}; HERE=$(cd $(dirname $0) && pwd)
TMP=$(mktemp -d)
( cd $TMP && cuobjdump -xelf all $HERE/../biped_mpc_py_b200/csrc/libbiped_mpc_b200.so >/dev/null && nvdisasm --print-line-info *.cubin > sass.txt )
ncu -i $REP --page source --csv --launch-skip $IDX --launch-count 1 2>/dev/null > $TMP/src.csv
python $HERE/ncu_by_function.py $TAG $TMP/src.csv $TMP/sass.txt
rm -rf $TMP
