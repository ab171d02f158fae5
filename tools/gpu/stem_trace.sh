#!/bin/bash
# debug build: where the roles of the persistent stem wait (cycles per role of CTA 0)
SPK_STEM_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-bins 2 2>&1 | grep "stem_p" | head -12
