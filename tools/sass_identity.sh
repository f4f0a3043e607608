#!/bin/bash
# tools/sass_identity.sh <git-rev> — proves that the device code of the working tree is the device code of <git-rev>
# (no GPU needed): builds libvgpu.so of that revision in a temporary worktree with the same nvcc command line and compares
# the complete `cuobjdump -sass` dumps. Used after host-only refactors made when no GPU was available: the kernels that
# ran green on a B200 at <git-rev> are, instruction for instruction, the kernels of HEAD.
set -e
rev=${1:?usage: tools/sass_identity.sh <git-rev>}
root=$(git rev-parse --show-toplevel)
tmp=$(mktemp -d /tmp/vgpu_sass_XXXXXX)
git -C "$root" worktree add --detach "$tmp/wt" "$rev" > /dev/null 2>&1
nvcc=${NVCC:-/usr/local/cuda/bin/nvcc}
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
$nvcc $flags -o "$tmp/base.so" "$tmp/wt/viyadb_b200/csrc/vgpu.cu" 2> /dev/null
$nvcc $flags -o "$tmp/head.so" "$root/viyadb_b200/csrc/vgpu.cu" 2> /dev/null
cuobjdump -sass "$tmp/base.so" | grep -v "^identifier = " > "$tmp/base.sass"   # the source path is the only thing that may differ
cuobjdump -sass "$tmp/head.so" | grep -v "^identifier = " > "$tmp/head.sass"
echo "# $(nvcc --version | tail -1)"
echo "# $rev = $(git -C "$root" rev-parse --short "$rev") ($(git -C "$root" log -1 --format=%s "$rev"))"
echo "# working tree = $(git -C "$root" rev-parse --short HEAD)$(git -C "$root" diff --quiet || echo ' + uncommitted changes')"
echo "kernels: $(grep -c 'Function :' "$tmp/base.sass") at $rev, $(grep -c 'Function :' "$tmp/head.sass") in the working tree"
echo "SASS lines: $(wc -l < "$tmp/base.sass") / $(wc -l < "$tmp/head.sass")"
echo "md5 $rev:          $(md5sum < "$tmp/base.sass" | cut -d' ' -f1)"
echo "md5 working tree:  $(md5sum < "$tmp/head.sass" | cut -d' ' -f1)"
if cmp -s "$tmp/base.sass" "$tmp/head.sass"; then echo "IDENTICAL: every kernel of the working tree is byte for byte the kernel of $rev"; rc=0
else echo "DIFFERENT:"; diff "$tmp/base.sass" "$tmp/head.sass" | head -40; rc=1; fi
git -C "$root" worktree remove --force "$tmp/wt" > /dev/null 2>&1
rm -rf "$tmp"
exit $rc
