#!/bin/bash
# usage: scripts/build_variant.sh NAME [-DFLAG=...]...   -> build/variants/NAME.so (same flags as asif_b200/_build.py plus the -D's)
set -e
name=$1; shift
out=build/variants; mkdir -p $out/obj_$name
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off"
nvcc $F -fmad=false "$@" -c -o $out/obj_$name/engine.o asif_b200/csrc/engine.cu &
nvcc $F -fmad=false "$@" -c -o $out/obj_$name/closed_loop.o asif_b200/csrc/closed_loop.cu &
nvcc $F -fmad=true "$@" -c -o $out/obj_$name/kernels_contract.o asif_b200/csrc/kernels_contract.cu &
nvcc $F -fmad=false "$@" -c -o $out/obj_$name/group.o asif_b200/csrc/group.cu &
if [ -f asif_b200/csrc/_obj/qp_admm.o ] && [ asif_b200/csrc/_obj/qp_admm.o -nt asif_b200/csrc/qp_admm.cuh ]; then cp asif_b200/csrc/_obj/qp_admm.o $out/obj_$name/qp_admm.o; else nvcc $F -fmad=true "$@" -c -o $out/obj_$name/qp_admm.o asif_b200/csrc/qp_admm.cu & fi
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
echo built $out/$name.so
