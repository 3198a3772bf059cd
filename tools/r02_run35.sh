#!/bin/bash
timeout 600 python -m pytest tests/test_multi_rank.py -q -m gpu 2>&1 | tail -5
