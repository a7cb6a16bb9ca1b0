#!/bin/bash
# Build radiation_ppo_b200/_C/var_<name>.so from the CURRENT tree with the step-kernel sources (rs_kernels.cu, rs_step1.cuh)
# of another commit: A/B timing of kernel revisions on the GPU box in one gpurun call (tools/gpu_variants.sh).
#   bash tools/build_variant_from.sh <commit> <name> [-DNAME=VALUE ...]
set -e
commit=$1; name=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/radiation_ppo_b200 $tmp/include
cp -r $root/radiation_ppo_b200/csrc $tmp/radiation_ppo_b200/
cp $root/include/*.h $tmp/include/
for f in rs_kernels.cu rs_step1.cuh; do git -C $root show $commit:radiation_ppo_b200/csrc/$f > $tmp/radiation_ppo_b200/csrc/$f; done
srcs=""; for s in rs_kernels.cu rs_gae.cu rs_maps.cu rs_pack.cu rs_rollout.cu; do srcs="$srcs $tmp/radiation_ppo_b200/csrc/$s"; done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -cudart static \
    "$@" -o $root/radiation_ppo_b200/_C/var_$name.so $srcs
rm -rf $tmp
echo built var_$name.so from $commit
