#!/bin/bash
# A/B of two builds of the library (ab/libA.so = baseline, in-tree = candidate): tokenizer time at 1 and 2 CTAs per SM
L=computer-vision-shoplifting-detection_b200/shopformer_b200/libshopformer_b200.so
cp $L ab/libCand.so
for v in A Cand; do
  cp ab/lib$v.so $L
  for occ in 2 3; do
    echo -n "build $v occ $occ: "; SF_TOK_OCC=$occ python profiles/kernel_times.py | tail -1
  done
done
cp ab/libCand.so $L
