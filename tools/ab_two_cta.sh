#!/bin/bash
# A/B of the two-CTAs-per-SM variant of the all-DMMA fp64 kernel (GCL_K1_MMA=4: 128 threads, 32-point tiles) against the default, with the
# second CTA of every SM started GCL_K1_STAGGER_NS later
export GCL_AB_PREC=fp64
echo "default:"; python tools/ab_k1.py 1M
for st in 0 900 2600 3500 5000; do echo "MMA=4 ctas=296 stagger=$st:"; GCL_K1_STAGGER_NS=$st GCL_K1_MMA=4 GCL_AB_CTAS=296 python tools/ab_k1.py 1M; done
