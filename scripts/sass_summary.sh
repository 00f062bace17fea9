#!/bin/bash
# SASS evidence for profiles/: instruction-count summary of the shipped libvqb_b200.so (tcgen05 / TMEM / TMA mnemonics per kernel)
so=${1:-multi-source-lms-for-audio_b200/libvqb_b200.so}
echo "# cuobjdump -sass $so  (sha256 $(sha256sum $so | cut -c1-16), $(stat -c %s $so) bytes)"
echo "# ldd: $(ldd $so | grep -E 'cudart|cublas|cudnn|nccl' | awk '{print $1}' | tr '\n' ' ')"
echo "## whole library"
cuobjdump -sass $so | grep -oE "\b(UTC[A-Z]*MMA(\.2CTA)?|LDTM(\.[xX0-9]+)*|STTM|UTMALDG\.[0-9]D(\.2CTA)?(\.MULTICAST)?|UTMASTG\.[0-9]D|UBLKCP[.A-Z]*|UBLKRED[.A-Z0-9]*|UTCBAR(\.2CTA)?(\.MULTICAST)?|UTCATOMSWS[.A-Z_]*|HMMA[.A-Z0-9]*|RED\.E\.ADD\.F32[.A-Za-z0-9]*|REDG?\.[A-Z0-9.]*|SYNCS[.A-Z0-9_]*|ELECT|FMNMX3?)\b" | sort | uniq -c | sort -rn
echo "## per kernel (demangled prefix): UTC*MMA / LDTM / UTMALDG / total instructions"
cuobjdump -sass $so | awk '
/Function : /{ if (name != "") printf "mma=%-3d ldtm=%-3d tma=%-3d instr=%-6d %s\n", mma, ldtm, tma, n, name; name=$3; mma=0; ldtm=0; tma=0; n=0; next }
/^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]+\*\/[ \t]+[A-Z@!]/{ n++; if ($0 ~ /UTC[A-Z]*MMA/) mma++; if ($0 ~ /LDTM/) ldtm++; if ($0 ~ /UTMALDG|UBLKCP/) tma++ }
END{ if (name != "") printf "mma=%-3d ldtm=%-3d tma=%-3d instr=%-6d %s\n", mma, ldtm, tma, n, name }' | c++filt | sed -E 's/\(CUtensorMap_st.*//' | sort -k5
