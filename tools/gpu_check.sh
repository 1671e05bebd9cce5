#!/bin/bash
# Runs the bring-up battery section by section, each in its own process under a timeout.
# usage: tools/gpu_check.sh [sections...]   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env, args...
  local name=$1; shift
  local envs=$1; shift
  echo "######## $name"
  env $envs timeout 400 python tools/gpu_check.py "$@" --out gpurun_out/check_$name.json > gpurun_out/check_$name.log 2>&1
  echo "exit=$?" >> gpurun_out/check_$name.log
  grep -E "^(conv|loss|model|full)/|^exit=|tc_available" gpurun_out/check_$name.log | cut -c1-220
}
SECTIONS=${@:-"losses conv_fp32 conv_mixed_tc models_fp32 models_mixed_tc full_mixed"}
for s in $SECTIONS; do
  case $s in
    losses) run losses "A=1" losses ;;
    conv_fp32) run conv_fp32 "DISPNET_B200_BACKEND=generic" conv --precision fp32 ;;
    conv_fp16_generic) run conv_fp16_generic "DISPNET_B200_BACKEND=generic" conv --precision fp16 ;;
    conv_fp16_tc) run conv_fp16_tc "A=1" conv --precision fp16 ;;
    conv_bf16_tc) run conv_bf16_tc "A=1" conv --precision bf16 ;;
    conv_mixed_tc) run conv_mixed_tc "A=1" conv --precision mixed ;;
    models_fp32) run models_fp32 "DISPNET_B200_BACKEND=generic" models --precision fp32 ;;
    models_fp16_generic) run models_fp16_generic "DISPNET_B200_BACKEND=generic" models --precision fp16 ;;
    models_fp16_tc) run models_fp16_tc "A=1" models --precision fp16 ;;
    models_mixed_tc) run models_mixed_tc "A=1" models --precision mixed ;;
    full_fp16) run full_fp16 "A=1" full --precision fp16 ;;
    full_fp32) run full_fp32 "DISPNET_B200_BACKEND=generic" full --precision fp32 ;;
    full_bf16) run full_bf16 "A=1" full --precision bf16 ;;
    full_mixed) run full_mixed "A=1" full --precision mixed ;;
    conv_tc32) run conv_tc32 "A=1" conv --precision tc32 ;;
    models_tc32) run models_tc32 "A=1" models --precision tc32 ;;
    full_tc32) run full_tc32 "A=1" full --precision tc32 ;;
  esac
done
