#!/bin/bash
# GroupNorm path comparison on the config-2 levels (native BFHWC layout): the default dispatch (slab kernel where a (domain, slab)
# fits 96 KB, slice ring otherwise) against each alternative forced through the environment.
run() { tag=$1; shift; echo "== $tag"; env "$@" python scripts/microbench.py --quick --only gn --iters 10 --out /dev/null 2>&1 | grep bfhwc | cut -c1-170; }
run "default (slab / ring)" CA_X=1
run "ring only" CA_GN_SLAB=0
run "slab with clusters of 8" CA_GN_SLAB_CLUSTER=8
run "streaming pair" CA_GN_SLAB=0 CA_GN_STREAM=1
run "team kernel" CA_GN_SLAB=0 CA_GN_RING=0
run "split launches" CA_GN_SLAB=0 CA_GN_RING=0 CA_GN_TEAM=0
