#!/usr/bin/env bash
cd tools/probe
for a in "8 6 64 80 0 4" "8 6 64 80 0 8" "8 6 64 80 0 5" "8 6 64 80 64000 76" "12 6 64 80 0 4" "24 24 32 80 0 24"; do timeout 60 ./tma_box_probe $a; done
