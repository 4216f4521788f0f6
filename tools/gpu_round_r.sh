#!/bin/bash
set -u
python tools/profile_encoder.py 2>&1 | tail -22
