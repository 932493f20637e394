#!/bin/bash
# The reference's own BSGSD.md example on one B200: server with -k 4096 (tables resident in HBM), puzzle 63 request.
# BSGSD.md quotes ~8 s for this request on a 64 GB / 8-thread CPU server.   usage: tools/bsgsd_demo.sh [k]
cd "${GRAFT_REPO_ROOT:-.}"
K=${1:-4096}
mkdir -p gpurun_out /tmp/bsgsd_demo && cd /tmp/bsgsd_demo
t0=$(date +%s.%N)
"$OLDPWD"/keyhunt_b200/keyhunt-b200-bsgsd -k $K -t 1 -p 18090 > server.log 2>&1 &
pid=$!
for i in $(seq 1 600); do grep -q "Listening in" server.log && break; kill -0 $pid 2>/dev/null || break; sleep 0.1; done
t1=$(date +%s.%N)
python - <<PY
import socket, time, json
def ask(line):
    t = time.time()
    s = socket.create_connection(("127.0.0.1", 18090), timeout=300)
    s.sendall(line)
    d = b""
    while True:
        x = s.recv(4096)
        if not x: break
        d += x
    return d.decode().strip(), time.time() - t
pk63 = b"0365ec2994b8cc0a20d40dd69edfe55ca32a54bcbbaa6b0ddcff36049301a54579"
out = {"k": $K, "startup_s": round($t1 - $t0, 2), "requests": []}
for rng in (b"4000000000000000:8000000000000000", b"4000000000000000:8000000000000000", b"1:4000000000000000"):
    r, dt = ask(pk63 + b" " + rng + b"\n")
    out["requests"].append({"range": rng.decode(), "reply": r, "seconds": round(dt, 4)})
print(json.dumps(out))
open("$OLDPWD/gpurun_out/bsgsd_demo_k$K.json", "w").write(json.dumps(out) + "\n")
PY
kill $pid; wait $pid 2>/dev/null
tail -5 server.log
