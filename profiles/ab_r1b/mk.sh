#!/bin/bash
# usage: mk.sh name [gitref]  -> creates /tmp/var/name with csrc+include from gitref (default be8033b)
name=$1; ref=${2:-HEAD}
rm -rf /tmp/var/$name; mkdir -p /tmp/var/$name/whisper_medusa_b200/csrc /tmp/var/$name/include; ln -s whisper_medusa_b200/csrc /tmp/var/$name/csrc
cd /root/repo
for f in $(git ls-tree --name-only $ref whisper_medusa_b200/csrc/); do git show $ref:$f > /tmp/var/$name/csrc/$(basename $f); done
git show $ref:include/whisper_medusa_b200.h > /tmp/var/$name/include/whisper_medusa_b200.h
