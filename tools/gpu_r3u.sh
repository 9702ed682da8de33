#!/bin/bash
for v in 2 1; do echo "=== V=$v"; SGG_CONV_V=$v timeout 200 python tools/conv_stack_events.py 2>&1 | tail -2; done
