#!/bin/bash
# GEMM pipeline depth experiment (debug library: TSSEP_GEMM_BN overrides the N tile; smaller tiles -> more stages)
for bn in 0 256 192 160 128; do
  echo "## TSSEP_GEMM_BN=$bn"
  TSSEP_DEBUG_KNOBS=1 TSSEP_GEMM_BN=$bn timeout 300 python scripts/profile_gemm.py --meetings 8 2>&1 | grep -v Warn | grep -E "pre_in|b0_in|b1_in|b2_in"
done
