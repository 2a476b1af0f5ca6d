#!/bin/bash
# three-part launch plans for the full 1500-model grid (T=200k statistics): top groups on wider clusters
mkdir -p gpurun_out
DIAG_TAG=pl DIAG_PLANS="4x2@0.3,0x0;4x4@0.032,4x2@0.268,0x0;4x4@0.064,4x2@0.236,0x0;2x4@0.016,4x2@0.284,0x0;2x4@0.032,4x2@0.268,0x0;4x4@0.032,4x2@0.2,0x0;4x4@0.064,4x2@0.17,0x0;4x2@0.25,0x0;4x2@0.35,0x0;4x2@0.2,0x0;4x4@0.1,4x2@0.2,0x0;2x4@0.016,4x4@0.05,4x2@0.2,0x0;4x4@0.05,4x2@0.15,4x1@0.15,0x0;4x2@0.2,4x1@0.2,0x0" timeout 900 python scripts/cd_dump.py 2>&1 | tee gpurun_out/r2n_cd_plans.log
