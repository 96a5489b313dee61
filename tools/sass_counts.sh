#!/bin/bash
# sass_counts.sh -- per-kernel counts of the Blackwell instructions in the built library (no GPU needed):
#   UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier
LIB=${1:-nmfgpu_b200/lib/libnmfgpu64.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ { name=$3; next }
{ for (i=1;i<=NF;i++) { if ($i ~ /^UTCHMMA/) a[name]++; if ($i ~ /^UTMALDG/) b[name]++; if ($i ~ /^LDTM/) c[name]++; if ($i ~ /^STTM/) d[name]++; if ($i ~ /^UTCBAR/) e[name]++; if ($i ~ /^SYNCS/) f[name]++; if ($i ~ /^FFMA/) g[name]++; if ($i ~ /^HMMA|^IMMA/) h[name]++; n[name]++ } }
END { printf "%-9s %-8s %-6s %-6s %-7s %-6s %-7s %s\n", "UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA", "kernel"; for (k in n) if (a[k]+b[k]+c[k]+d[k]+g[k] > 0) printf "%-9d %-8d %-6d %-6d %-7d %-6d %-7d %s\n", a[k], b[k], c[k], d[k], e[k], f[k], g[k], k }' | sort -k8 | c++filt | cut -c1-200
