#!/bin/bash
# usage: tools/exp_build.sh <name>     snapshot of the CUDA sources -> an experiment library biped_mpc_py_b200/csrc/_exp/lib_<name>.so
# (h = 10 lane unit rebuilt from the snapshot, the other objects taken from the last full build); tools/lane_probe.py loads it
# with EXP_LIB=<name>.  For A/B measurements of kernel variants in one GPU call.
set -e
name=$1; root=$(cd $(dirname $0)/.. && pwd); d=/tmp/exp/$name
rm -rf $d; mkdir -p $d/biped_mpc_py_b200 $root/biped_mpc_py_b200/csrc/_exp
cp -r $root/include $d/include
mkdir -p $d/biped_mpc_py_b200/csrc && cp $root/biped_mpc_py_b200/csrc/*.cu $root/biped_mpc_py_b200/csrc/*.cuh $root/biped_mpc_py_b200/csrc/*.h $d/biped_mpc_py_b200/csrc/
cd $d/biped_mpc_py_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DBMPC_LANE_UNIT=10 $EXP_FLAGS -c -o lane_h10.o bmpc_lane.cu 2>&1 | grep -v deprecated || true
if [ "$2" = "all" ]; then
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -c -o bmpc.o bmpc.cu 2>&1 | grep -v deprecated || true
else
  cp $root/biped_mpc_py_b200/csrc/_obj/bmpc.o bmpc.o
fi
nvcc -shared -o $root/biped_mpc_py_b200/csrc/_exp/lib_$name.so bmpc.o lane_h10.o $root/biped_mpc_py_b200/csrc/_obj/bmpc_lane_h30.o 2>&1 | grep -v deprecated || true
echo "built $name"
