#!/bin/bash
# tests (-k expr) then bench variants.  usage: bash profiles/run_tv.sh "<-k expr>" "ENV=.." ...
K="$1"; shift
bash profiles/run_tests.sh "$K" | tail -15
bash profiles/run_variants.sh "$@"
