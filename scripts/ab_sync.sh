#!/bin/bash
# A/B of the lagged active-count readback against the blocking one (2048 instances)
SKIP_TESTS=1 bash scripts/pcg_ab.sh
echo "--- B2T_SYNC_PASSES=1"
SKIP_TESTS=1 B2T_SYNC_PASSES=1 bash scripts/pcg_ab.sh
