#!/bin/bash
# compute-sanitizer pass over the kernels: memcheck, racecheck (shared-memory hazards between the warp roles),
# synccheck (barrier misuse) on the kernel-level tests and on one CUDA-graph decode replay. Run under gpurun:
#   bash scripts/sanitize.sh        -> gpurun_out/sanitize_<tool>.log + gpurun_out/sanitize_summary.txt
# The sanitizer serialises kernels and slows them 10-100x: the big shapes are deselected (-k), mbarrier watchdogs stay armed.
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_kernels_gpu.py tests/test_gemv_gpu.py"
SEL='not (2056 or 22016 or 11008 or 32000 or 12304 or stream_k or full_width or bench_shape)'
: > gpurun_out/sanitize_summary.txt
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize_${tool}.log
  MYR_SANITIZE=1 timeout ${T_SAN:-900} $SAN --tool $tool --print-limit 20 --error-exitcode 86 \
      python -m pytest $TESTS -m gpu -q -x -k "$SEL" > $log 2>&1
  rc=$?
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
  echo "$tool: exit $rc | ${errs:-no summary line} | $(grep -E 'passed|failed' $log | tail -1)" >> gpurun_out/sanitize_summary.txt
done
# one decode-graph replay (greedy_decode captures the step, then replays it) under memcheck
MYR_SANITIZE=1 timeout ${T_SAN:-900} $SAN --tool memcheck --print-limit 20 --error-exitcode 86 \
    python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "llama_tiny_logits_and_greedy" > gpurun_out/sanitize_decode_graph.log 2>&1
rc_graph=$?
# the vision-expert heads (csrc/expert.cu) and its trunk on the tiny configuration
MYR_SANITIZE=1 timeout ${T_SAN:-900} $SAN --tool memcheck --print-limit 20 --error-exitcode 86 \
    python -m pytest tests/test_expert_gpu.py -m gpu -q -x -k "not full_width" > gpurun_out/sanitize_expert.log 2>&1
echo "memcheck vision expert: exit $? | $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_expert.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize_expert.log | tail -1)" >> gpurun_out/sanitize_summary.txt
# the training-side kernels (fused attention backward, fused LoRA kernels, conv / adaptor reductions) under memcheck and racecheck
for tool in memcheck racecheck; do
  MYR_SANITIZE=1 timeout ${T_SAN:-900} $SAN --tool $tool --print-limit 20 --error-exitcode 86 \
      python -m pytest tests/test_training_gpu.py -m gpu -q -x -k "attention_bwd or lora_dropout or conv_trunk or norm_bwd or swiglu_gelu_rope or clamp_ce" \
      > gpurun_out/sanitize_train_${tool}.log 2>&1
  echo "$tool training kernels: exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_train_${tool}.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize_train_${tool}.log | tail -1)" >> gpurun_out/sanitize_summary.txt
done
echo "memcheck decode graph: exit $rc_graph | $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_decode_graph.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize_decode_graph.log | tail -1)" >> gpurun_out/sanitize_summary.txt
cat gpurun_out/sanitize_summary.txt
