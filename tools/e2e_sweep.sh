for cfg in "32768 8192 2048" "65536 8192 2048" "65536 16384 2048" "65536 16384 16384" "65536 32768 32768" "65536 10944 10944" "65536 21888 21888"; do
  set -- $cfg
  echo -n "pass=$1 chunk=$2 first=$3: "
  MDC_VT_PASS=$1 MDC_VT_CHUNK=$2 MDC_VT_FIRST=$3 python tools/e2e_probe.py | grep -E "probs\+hist|stream" | awk '{printf "%s %s ms | ", $2, $(NF-3)}'; echo
done
