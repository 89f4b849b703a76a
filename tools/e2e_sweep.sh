for cfg in "32768 8192 2048" "32768 16384 4096" "32768 4096 2048" "65536 8192 2048" "65536 16384 4096" "21888 8192 2048" "32768 10944 2048"; do
  set -- $cfg
  echo -n "pass=$1 chunk=$2 first=$3: "
  MDC_VT_PASS=$1 MDC_VT_CHUNK=$2 MDC_VT_FIRST=$3 python tools/e2e_probe.py | grep "probs+hist"
done
