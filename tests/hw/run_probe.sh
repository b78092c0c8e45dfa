#!/bin/bash
# runs every probe of tests/hw/umma_probe.cu in its own process (B200 only)
cd "$(dirname "$0")/_build" || exit 1
for t in ${PROBES:-1 2 3 4 5 6 7 8 9 10 11}; do timeout 60 ./umma_probe $t || echo "test $t: exit $?"; done
