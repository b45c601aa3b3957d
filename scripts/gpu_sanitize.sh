#!/bin/bash
# compute-sanitizer over the smoke configuration (E=64, N=10, 100 steps + oracle check) for both step kernels and the post
# kernel: memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards; the pf kernel has five mbarrier
# families and a proxy-fence protocol), synccheck.  Logs -> gpurun_out/sanitizer_<tool>_<kernel>.log
for kern in pf generic; do
  for tool in memcheck racecheck synccheck; do
    FLEETSTEP_KERNEL=$kern timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke \
        > gpurun_out/sanitizer_${tool}_${kern}.log 2>&1
    echo "== $tool $kern: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_${tool}_${kern}.log | tr '\n' ' ')"
  done
done
