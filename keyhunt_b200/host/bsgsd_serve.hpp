// bsgsd_serve.hpp — the BSGS server front end: the reference's `bsgsd` request surface
// (bsgsd.cpp:3307 client_handler, BSGSD.md) in front of BSGS tables that stay resident in HBM.
//
// Included by keyhunt_gpu.cpp when it is compiled as keyhunt-b200-bsgsd (-DKH_BSGSD); uses its U256 / parsing helpers.
//
//   line protocol   "<pubkey hex> <from hex>:<to hex>\n"      (also "<pubkey> <from> <to>")
//                   -> "<private key hex>\n" | "404 Not Found\n" | "400 Bad Request"
//   HTTP            POST with a JSON body {"pubkey": "...", "from": "...", "to": "..."}
//                   -> 200 / 404 / 400 with the same body text, Content-Length, Connection: close, X-Elapsed-Seconds
// The search is the sequential one (thread_process_bsgs, windows of 2N keys from `from` while below `to`), split
// over the GPUs in contiguous blocks of windows.  Requests are served one at a time: the reference starts a thread
// per client but keeps the request in process-wide globals (n_range_start, BSGS_CURRENT, bsgs_found), so concurrent
// requests are not something its clients can rely on either.
//
// Deliberate differences (both are crashes or footguns of the reference, not behaviour a client can use):
//  * a malformed public key gets "400 Bad Request" (the reference's ParsePublicKeyHex exits the whole server);
//  * the default listen address is 127.0.0.1 as BSGSD.md documents (the reference binary binds 0.0.0.0);
//  * the table files are read / written only with -S (the reference server always does): rebuilding in HBM is faster
//    than reading them back.
#pragma once
#include <arpa/inet.h>
#include <netinet/in.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <chrono>

static volatile sig_atomic_t bsgsd_stop = 0;
static int bsgsd_listen_fd = -1;
static void bsgsd_on_signal(int) { bsgsd_stop = 1; if (bsgsd_listen_fd >= 0) shutdown(bsgsd_listen_fd, SHUT_RDWR); }

static bool bsgsd_send(int fd, const std::string &s) {
  size_t sent = 0;
  while (sent < s.size()) {
    ssize_t n = send(fd, s.data() + sent, s.size() - sent, MSG_NOSIGNAL);
    if (n <= 0) return false;
    sent += (size_t)n;
  }
  return true;
}

struct BsgsdRequest { bool http = false; std::string pubkey, from, to; };

// 1 = parsed, 0 = malformed (caller answers 400), -1 = connection gone (no answer), -2 = too large (HTTP 413)
// A request must arrive within BSGSD_READ_DEADLINE_S seconds IN TOTAL: the per-recv timeout alone would let a client that
// trickles one byte every few seconds hold the (one-request-at-a-time) server for weeks.
#ifndef BSGSD_READ_DEADLINE_S
#define BSGSD_READ_DEADLINE_S 15
#endif
static ssize_t bsgsd_recv(int fd, char *buf, size_t cap, int flags, const std::chrono::steady_clock::time_point &t0) {
  const double left = BSGSD_READ_DEADLINE_S - std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (left <= 0) return -1;
  struct timeval tv = {(time_t)left, (suseconds_t)((left - (double)(time_t)left) * 1e6) + 1};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  return recv(fd, buf, cap, flags);
}
#define recv_deadline(fd, buf, cap, flags) bsgsd_recv(fd, buf, cap, flags, t_start)
static int bsgsd_read_request(int fd, BsgsdRequest &rq) {
  const auto t_start = std::chrono::steady_clock::now();
  char buf[1024];
  ssize_t n = recv_deadline(fd, buf, sizeof(buf) - 1, MSG_PEEK);
  if (n <= 0) return -1;
  rq.http = (memcmp(buf, "POST", (size_t)std::min<ssize_t>(n, 4)) == 0);
  if (rq.http) {
    std::string req;
    size_t hdr_end;
    do {
      n = recv_deadline(fd, buf, sizeof(buf), 0);
      if (n <= 0) return -1;
      req.append(buf, (size_t)n);
      if (req.size() > (1u << 20)) return -2;
      hdr_end = req.find("\r\n\r\n");
    } while (hdr_end == std::string::npos);
    const std::string head = req.substr(0, hdr_end);
    std::string body = req.substr(hdr_end + 4);
    size_t content_length = 0, p = head.find("Content-Length:");
    if (p != std::string::npos) {
      p += 15;
      while (p < head.size() && (head[p] == ' ' || head[p] == '\t')) p++;
      content_length = strtoull(head.c_str() + p, NULL, 10);
    }
    while (body.size() < content_length) {
      n = recv_deadline(fd, buf, sizeof(buf), 0);
      if (n <= 0) return -1;
      body.append(buf, (size_t)n);
      if (body.size() > (1u << 20)) return -2;
    }
    auto value_of = [&](const char *key, std::string &out) {   // "key" ... : ... "value"
      const std::string needle = std::string("\"") + key + "\"";
      size_t q = body.find(needle);
      if (q == std::string::npos) return false;
      q = body.find(':', q + needle.size());
      if (q == std::string::npos) return false;
      q = body.find('"', q);
      if (q == std::string::npos) return false;
      const size_t e = body.find('"', q + 1);
      if (e == std::string::npos) return false;
      out = body.substr(q + 1, e - q - 1);
      return true;
    };
    return (value_of("pubkey", rq.pubkey) && value_of("from", rq.from) && value_of("to", rq.to)) ? 1 : 0;
  }
  std::string line;
  do {
    n = recv_deadline(fd, buf, sizeof(buf), 0);
    if (n <= 0) return -1;
    line.append(buf, (size_t)n);
    if (line.size() > 4096) return 0;
  } while (line.find('\n') == std::string::npos);
  // stringtokenizer (util.c:245): trim "\t\n\r :" at both ends, split on " \t:"
  const size_t a = line.find_first_not_of("\t\n\r :"), b = line.find_last_not_of("\t\n\r :");
  std::vector<std::string> tok;
  if (a != std::string::npos) {
    const std::string t = line.substr(a, b - a + 1);
    size_t i = 0;
    while (i < t.size()) {
      const size_t s = t.find_first_not_of(" \t:", i);
      if (s == std::string::npos) break;
      size_t e = t.find_first_of(" \t:", s);
      if (e == std::string::npos) e = t.size();
      tok.push_back(t.substr(s, e - s));
      i = e;
    }
  }
  if (tok.size() < 3) return 0;   // two tokens can never hold a "from:to" pair once ':' is a separator (bsgsd.cpp:3447)
  rq.pubkey = tok[0]; rq.from = tok[1]; rq.to = tok[2];
  return 1;
}

static bool bsgsd_parse_pubkey(const std::string &s, uint8_t xy[64], bool &compressed) {   // ParsePublicKeyHex SECP256K1.cpp:270
  uint8_t raw[65];
  if (s.size() == 66 && is_hex(s) && hex2bin(s, raw, 33) && (raw[0] == 2 || raw[0] == 3)) {
    if (!decompress_pub(raw + 1, raw[0] & 1, xy + 32)) return false;
    memcpy(xy, raw + 1, 32); compressed = true;
    return true;
  }
  if (s.size() == 130 && is_hex(s) && hex2bin(s, raw, 65) && raw[0] == 4) {
    // the point must be on the curve with canonical coordinates (x, y < p, y^2 = x^3 + 7): the y the curve gives for this x and
    // this parity must be the y that was sent; anything else gets 400 like a malformed compressed key
    uint8_t y[32];
    if (!decompress_pub(raw + 1, raw[64] & 1, y) || memcmp(y, raw + 33, 32) != 0) return false;
    memcpy(xy, raw + 1, 64); compressed = false;
    return true;
  }
  return false;
}
static bool bsgsd_hex_ok(const std::string &s) { for (char c : s) if (!isxdigit((unsigned char)c)) return false; return s.size() <= 64; }   // isValidHex util.c:347

// sequential search of [from, to) over all GPUs: contiguous blocks of 2N-key windows per GPU
static int bsgsd_search(std::vector<kh_ctx *> &gpus, const kh_bsgs_desc &d, const uint8_t xy[64], const U256 &from, const U256 &to, U256 &key) {
  if (u_cmp(from, to) >= 0) return 0;
  const U256 two_n = u_mul_u64(u_from_u64(d.n), 2);
  const U256 width = u_sub(to, from);
  for (int i = 0; i < 16; i++) if (width.b[i]) return -1;                       // library limit: 2^128 keys per request
  unsigned __int128 w = 0;
  for (int i = 16; i < 32; i++) w = (w << 8) | width.b[i];
  const unsigned __int128 step = (unsigned __int128)2 * d.n, nw = (w + step - 1) / step;
  if (nw > ((unsigned __int128)1 << 50)) return -1;
  const uint64_t windows = (uint64_t)nw, ng = gpus.size();
  std::vector<int> fnd(ng, 0), err(ng, 0);
  std::vector<U256> keys(ng);
  std::atomic<bool> stop_all{false};
  auto worker = [&](size_t g) {
    const uint64_t base = windows / ng, rem = windows % ng;
    const uint64_t first = g * base + std::min<uint64_t>(g, rem), cnt = base + (g < rem ? 1 : 0);
    if (!cnt) return;
    const U256 f = u_add(from, u_mul_u64(two_n, first));
    U256 t = u_add(f, u_mul_u64(two_n, cnt));
    if (first + cnt == windows) t = to;                                          // the last window keeps the reference's overshoot past `to`
    // in slices of 2^24 windows (a few seconds of giant steps) so that SIGINT / SIGTERM can end a huge request between slices
    const uint64_t slice = 1ULL << 24;
    for (uint64_t done = 0; done < cnt && !fnd[g] && !err[g] && !stop_all.load(); done += slice) {
      if (bsgsd_stop) { err[g] = 1; break; }
      const uint64_t c = std::min<uint64_t>(slice, cnt - done);
      const U256 sf = u_add(f, u_mul_u64(two_n, done));
      const U256 st = (done + c == cnt) ? t : u_add(sf, u_mul_u64(two_n, c));
      if (kh_bsgs_search(gpus[g], xy, sf.b, st.b, keys[g].b, &fnd[g]) != KH_OK) { fprintf(stderr, "[E] %s\n", kh_last_error(gpus[g])); err[g] = 1; }
      if (fnd[g]) stop_all.store(true);                                         // first key found ends the request on every GPU
    }
  };
  std::vector<std::thread> th;
  for (size_t g = 0; g < ng; g++) th.emplace_back(worker, g);
  for (auto &t : th) t.join();
  for (size_t g = 0; g < ng; g++) if (err[g]) return -1;
  for (size_t g = 0; g < ng; g++) if (fnd[g]) { key = keys[g]; return 1; }
  return 0;
}

static int bsgsd_serve(std::vector<kh_ctx *> &gpus, const kh_bsgs_desc &d, const char *ip, int port) {
  int srv = socket(AF_INET, SOCK_STREAM, 0);
  if (srv < 0) { perror("socket failed"); return EXIT_FAILURE; }
  int opt = 1;
  setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &opt, sizeof(opt));
  struct sockaddr_in addr;
  memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  if (port <= 0 || port > 65535) { fprintf(stderr, "[W] Invalid port %d, defaulting to 8080\n", port); port = 8080; }
  if (!ip || !ip[0] || inet_pton(AF_INET, ip, &addr.sin_addr) != 1) { if (ip && ip[0]) fprintf(stderr, "[W] Invalid IP address: %s, defaulting to 127.0.0.1\n", ip); ip = "127.0.0.1"; inet_pton(AF_INET, ip, &addr.sin_addr); }
  addr.sin_port = htons((uint16_t)port);
  if (bind(srv, (struct sockaddr *)&addr, sizeof(addr)) < 0) { perror("bind failed"); return EXIT_FAILURE; }
  if (listen(srv, 16) < 0) { perror("listen failed"); return EXIT_FAILURE; }
  bsgsd_listen_fd = srv;
  signal(SIGINT, bsgsd_on_signal);
  signal(SIGTERM, bsgsd_on_signal);
  printf("[+] Listening in %s:%i\n", ip, port);
  fflush(stdout);
  while (!bsgsd_stop) {
    struct sockaddr_in peer;
    socklen_t plen = sizeof(peer);
    const int fd = accept(srv, (struct sockaddr *)&peer, &plen);
    if (fd < 0) { if (bsgsd_stop) break; perror("accept failed"); continue; }
    char pip[INET_ADDRSTRLEN];
    inet_ntop(AF_INET, &peer.sin_addr, pip, sizeof(pip));
    printf("[+] Accepting incoming conection from %s:%i\n", pip, ntohs(peer.sin_port));
    struct timeval tv = {10, 0};                                                  // a silent client must not park the server
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    const auto t0 = std::chrono::steady_clock::now();
    BsgsdRequest rq;
    const int pr = bsgsd_read_request(fd, rq);
    const char *bad = rq.http ? "HTTP/1.1 400 Bad Request\r\nConnection: close\r\n\r\n" : "400 Bad Request";
    uint8_t xy[64];
    bool compressed = false;
    U256 from, to, key;
    if (pr == -2) bsgsd_send(fd, "HTTP/1.1 413 Request Entity Too Large\r\nConnection: close\r\n\r\n");
    else if (pr == 0) { printf("Invalid input format from client\n"); bsgsd_send(fd, bad); }
    else if (pr == 1 && !bsgsd_parse_pubkey(rq.pubkey, xy, compressed)) { printf("Invalid publickey format from client %s\n", rq.pubkey.c_str()); bsgsd_send(fd, bad); }
    else if (pr == 1 && !(bsgsd_hex_ok(rq.from) && bsgsd_hex_ok(rq.to))) { printf("Invalid hexadecimal format from client %s:%s\n", rq.from.c_str(), rq.to.c_str()); bsgsd_send(fd, bad); }
    else if (pr == 1) {
      from = u_zero(); to = u_zero();
      if (!rq.from.empty()) u_from_hex(from, rq.from.c_str());
      if (!rq.to.empty()) u_from_hex(to, rq.to.c_str());
      const int r = bsgsd_search(gpus, d, xy, from, to, key);
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (r < 0) bsgsd_send(fd, bad);
      else {
        std::string body = (r == 1) ? (u_hex(key) + "\n") : std::string("404 Not Found\n");
        if (r == 1) {
          const std::string pubhex = compressed ? (std::string((xy[63] & 1) ? "03" : "02") + hex_of(xy, 32)) : ("04" + hex_of(xy, 64));
          printf("[+] Thread Key found privkey %s   \n[+] Publickey %s\n", u_hex(key).c_str(), pubhex.c_str());
        }
        if (rq.http) {
          char head[256];
          snprintf(head, sizeof(head), "%sContent-Type: text/plain\r\nContent-Length: %zu\r\nConnection: close\r\nX-Elapsed-Seconds: %.3f\r\n\r\n",
                   (r == 1) ? "HTTP/1.1 200 OK\r\n" : "HTTP/1.1 404 Not Found\r\n", body.size(), secs);
          body = std::string(head) + body;
        }
        if (!bsgsd_send(fd, body)) printf("Failed to send message to client\n");
      }
    }
    close(fd);
    printf("[+] Closing conection from %s:%i\n", pip, ntohs(peer.sin_port));
    fflush(stdout);
  }
  close(srv);
  return 0;
}
