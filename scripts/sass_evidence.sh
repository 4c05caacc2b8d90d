#!/usr/bin/env bash
# Blackwell-specific SASS mnemonics per kernel of the shipped library (see profiles/r02_sass_evidence.txt).
cuobjdump -sass "$(dirname "$0")/../reid-gan_b200/libreid_b200.so" | grep -E 'Function :|UTC[A-Z]*MMA|UTCBAR|UTMALDG|LDTM|DMMA'
