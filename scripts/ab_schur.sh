set -x
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
SKIP_TESTS=1 bash scripts/pcg_ab.sh
echo "--- V1"
SKIP_TESTS=1 B2T_SCHUR_V1=1 bash scripts/pcg_ab.sh
