#!/bin/bash
# rolling row pipeline in the panel of the cluster CD kernel: bit-identity, then timers under load
mkdir -p gpurun_out
for v in 8 9 10; do
  SGLM_CDC_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_cd_cluster.py -x -q -k "every_shape or multi_part or wide_design" > gpurun_out/r2o_pytest_v$v.log 2>&1; echo "variant $v pytest rc=$?"; tail -2 gpurun_out/r2o_pytest_v$v.log
done
DIAG_EVERY=1 DIAG_PLANS="4x2@0.3,0x0#0;4x2@0.3,0x0#8;4x2@0.3,0x0#9;4x2@0.3,0x0#10;4x2@0.3,0x0#11;4x4@0.032,4x2@0.268,0x0#0;4x4@0.032,4x2@0.268,0x0#9;4x4@0.032,4x2@0.268,0x0#10;4x4@0.3,0x0#10;4x4@0.3,0x0#9" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r2o_cd_timers.log
