#!/bin/bash
# A/B of several builds of the library: every ab/lib_*.so (git-ignored, travels with gpurun) takes the in-tree library's place
# in turn and profiles/kernel_times.py prints the kernel times.  usage: bash profiles/ab_libs.sh [config ...]
L=computer-vision-shoplifting-detection_b200/shopformer_b200/libshopformer_b200.so
cp $L ab/.saved.so
for f in ab/lib_*.so; do
  cp $f $L
  for cfg in ${@:-A}; do
    echo -n "$(basename $f) $cfg: "; python profiles/kernel_times.py $cfg | tail -1
  done
done
cp ab/.saved.so $L
