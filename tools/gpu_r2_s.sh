#!/usr/bin/env bash
timeout 120 python tools/probe/roi_cl_diag.py 2>&1 | tail -14
