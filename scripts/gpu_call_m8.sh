#!/bin/bash
# round 2, session 3, call 1: groups of 8 models — bit-identity tests, then timelines under several plans
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cd_cluster.py -x -q > gpurun_out/r2m_pytest_cd.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2m_pytest_cd.log
DIAG_TAG=m8 DIAG_PLANS="4x2@0.3,0x0;8x4@0.3,0x0;8x2@0.3,0x0;8x4@0.15,4x2@0.15,0x0;8x4@0.2,0x0;8x4@0.4,0x0" timeout 600 python scripts/cd_dump.py 2>&1 | tee gpurun_out/r2m_cd_dump.log
