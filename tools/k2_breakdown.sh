#!/bin/bash
# builds the timing-only variants tools/k2_breakdown.py runs (variants/ is git-ignored but travels to the GPU box)
B=tools/build_variant.sh
$B x0_base -DDEX_EXP &
$B x1_fix3 -DDEX_EXP -DDEX_EXP_FIXITER=3 &
$B x2_fix3_noeuler -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOEULER &
$B x3_fix3_nofactor -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOFACTOR &
wait
$B x4_fix3_noapply -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOAPPLY &
$B x5_fix3_nomatvec -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOMATVEC &
$B x6_fix3_nols -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOLS &
$B x7_fix3_none -DDEX_EXP -DDEX_EXP_FIXITER=3 -DDEX_EXP_NOEULER -DDEX_EXP_NOFACTOR -DDEX_EXP_NOAPPLY -DDEX_EXP_NOMATVEC -DDEX_EXP_NOLS &
wait
$B x8_fix0 -DDEX_EXP -DDEX_EXP_FIXITER=0 -DDEX_EXP_NOEULER &
$B x9_fix1 -DDEX_EXP -DDEX_EXP_FIXITER=1 &
wait
ls -la variants/
