cd "$GRAFT_REPO_ROOT"
export DIAG_SHAPES=4x2 DIAG_PLANS="4x2@0.3,0x0"
for v in 0 3 4; do
  echo "=== SGLM_CDC_VARIANT=$v"
  SGLM_CDC_VARIANT=$v timeout 300 python scripts/cd_cluster_check.py 2>&1 | grep -v "wu)\|warm-up" | tail -8
done
timeout 600 python -m pytest tests/test_gpu_at_scale.py -x -q -k "poisson or ols or holdout or preprocess or second" 2>&1 | tail -15
timeout 300 python scripts/config_bench.py c2 --no-cpu 2>&1 | cut -c1-700
