#!/bin/bash
# GroupNorm fused (single launch, cross-CTA exchange) vs split (stats / finalize / apply) comparison
for m in split fused; do echo "== CA_GN_MODE=$m"; CA_GN_MODE=$m python scripts/microbench.py --quick --iters 10 2>&1 | grep -E "groupnorm" | cut -c1-200; done
