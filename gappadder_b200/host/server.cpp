#include "server.hpp"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>

#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

namespace gpm {

namespace {

float parse_float(const char* s) { float v = 0; if (s) sscanf(s, "%f", &v); return v; }   // CM/main.cpp:91-93
int parse_int(const char* s, int dflt) { int v = dflt; if (s) sscanf(s, "%d", &v); return v; }

bool read_all(int fd, void* buf, size_t n)
{
    char* p = (char*)buf;
    while (n) {
        const ssize_t got = ::read(fd, p, n);
        if (got <= 0) return false;
        p += got; n -= (size_t)got;
    }
    return true;
}
bool write_all(int fd, const void* buf, size_t n)
{
    const char* p = (const char*)buf;
    while (n) {
        const ssize_t put = ::send(fd, p, n, MSG_NOSIGNAL);
        if (put <= 0) return false;
        p += put; n -= (size_t)put;
    }
    return true;
}
bool put_blob(int fd, const std::string& s)
{
    const uint32_t n = (uint32_t)s.size();
    return write_all(fd, &n, 4) && (n == 0 || write_all(fd, s.data(), n));
}
bool get_blob(int fd, std::string& s, uint32_t limit = 1u << 30)
{
    uint32_t n = 0;
    if (!read_all(fd, &n, 4) || n > limit) return false;
    s.resize(n);
    return n == 0 || read_all(fd, &s[0], n);
}

bool send_reply(int fd, const Reply& r)
{
    const int32_t code = r.exit_code;
    const uint32_t flags = r.wrote_info ? 1u : 0u;
    return write_all(fd, "GPR1", 4) && write_all(fd, &code, 4) && write_all(fd, &flags, 4) && put_blob(fd, r.out) && put_blob(fd, r.info) &&
           put_blob(fd, r.gml) && put_blob(fd, r.err);
}

// Requests are batched together only when every option that reaches merge_gaps is the same.
std::string options_key(const MergeOptions& o)
{
    char b[512];
    snprintf(b, sizeof b, "%a|%a|%a|%a|%a|%d|%d|%d|%d|%a|%a|%d|%d|%d|%d", o.max_frac_score_loss, o.min_frac_overlap, o.min_overlap_len,
             o.max_overlap_clip_len, o.min_overlap_len_with_scaffold, o.num_threads, o.min_support_kmer, o.line_length, o.quick_kmer_len,
             o.score_mismatch, o.score_indel, o.max_contig_path_len, o.max_count_contig_in_path, (int)o.host_quick_check, (int)o.host_relax);
    return b;
}

struct Pending {
    int fd;
    Request req;
};

std::atomic<bool> g_stop(false);
int g_listen_fd = -1;
void on_signal(int) { g_stop = true; if (g_listen_fd >= 0) ::shutdown(g_listen_fd, SHUT_RDWR); }

} // namespace

bool parse_request(int argc, const char* const* argv, Request& r)
{
    int pos = 1;
    bool have_input = false;
    while (pos < argc) {
        const char* a = argv[pos];
        const char* val = pos + 1 < argc ? argv[pos + 1] : nullptr;
        if (a[0] != '-') { r.input = a; have_input = true; ++pos; continue; }
        if (!strcmp(a, "--host-quick-check")) { r.opt.host_quick_check = true; ++pos; continue; }
        if (!strcmp(a, "--host-relax")) { r.opt.host_relax = true; ++pos; continue; }
        if (!strcmp(a, "--no-gml")) { r.write_gml = false; ++pos; continue; }
        switch (a[1]) {
        case 'V': r.opt.verbose = true; ++pos; break;
        case 'l': r.opt.line_length = parse_int(val, r.opt.line_length); pos += 2; break;
        case 's': r.opt.max_frac_score_loss = parse_float(val); pos += 2; break;
        case 'c': r.opt.min_frac_overlap = parse_float(val); pos += 2; break;
        case 'x': r.opt.min_overlap_len = parse_float(val); pos += 2; break;
        case 'y': r.opt.max_overlap_clip_len = parse_float(val); pos += 2; break;
        case 'm': r.opt.min_support_kmer = parse_int(val, r.opt.min_support_kmer); pos += 2; break;
        case 't': r.opt.num_threads = parse_int(val, r.opt.num_threads); pos += 2; break;
        case 'z': r.opt.min_overlap_len_with_scaffold = parse_float(val); pos += 2; break;
        case 'k': r.opt.quick_kmer_len = parse_int(val, r.opt.quick_kmer_len); pos += 2; break;
        case 'i':
            if (a[2] == '1') { r.opt.score_mismatch = parse_float(val); pos += 2; break; }
            if (a[2] == '2') { r.opt.score_indel = parse_float(val); pos += 2; break; }
            return false;
        case 'o': if (val) r.opt.info_file = val; pos += 2; break;
        case 'p':
            if (a[2] == '1') { r.opt.max_contig_path_len = parse_int(val, -1); pos += 2; break; }
            if (a[2] == '2') { r.opt.max_count_contig_in_path = parse_int(val, -1); pos += 2; break; }
            return false;
        case 'e': pos += 2; break;                         // scaffold info file: unused by CompactVer3
        case 'u': pos += 2; break;                         // support-pairs cutoff: unused by CompactVer3
        default: return false;
        }
    }
    if (!have_input && argc > 1) r.input = argv[1];        // argv[repeatfileArgIndex] defaults to argv[1] in the reference
    return true;
}

bool request_from_server(const std::string& socket_path, int argc, char** argv, Reply& reply)
{
    const int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return false;
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    if (socket_path.size() >= sizeof(addr.sun_path)) { ::close(fd); return false; }
    strcpy(addr.sun_path, socket_path.c_str());
    if (::connect(fd, (sockaddr*)&addr, sizeof addr) != 0) { ::close(fd); return false; }
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) cwd[0] = 0;
    const uint32_t n = (uint32_t)argc;
    bool ok = write_all(fd, "GPM1", 4) && write_all(fd, &n, 4);
    for (int i = 0; ok && i < argc; ++i) ok = put_blob(fd, argv[i]);
    ok = ok && put_blob(fd, cwd);
    char magic[4];
    int32_t code = 0;
    uint32_t flags = 0;
    ok = ok && read_all(fd, magic, 4) && !memcmp(magic, "GPR1", 4) && read_all(fd, &code, 4) && read_all(fd, &flags, 4) &&
         get_blob(fd, reply.out) && get_blob(fd, reply.info) && get_blob(fd, reply.gml) && get_blob(fd, reply.err);
    ::close(fd);
    if (!ok) return false;
    reply.exit_code = code;
    reply.wrote_info = (flags & 1u) != 0;
    return true;
}

int serve(const ServeOptions& so)
{
    ::unlink(so.socket_path.c_str());
    g_listen_fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
    if (g_listen_fd < 0) { perror("socket"); return 2; }
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    if (so.socket_path.size() >= sizeof(addr.sun_path)) { fprintf(stderr, "ContigsMerger_b200: socket path too long\n"); return 2; }
    strcpy(addr.sun_path, so.socket_path.c_str());
    if (::bind(g_listen_fd, (sockaddr*)&addr, sizeof addr) != 0 || ::listen(g_listen_fd, 1024) != 0) { perror("bind/listen"); return 2; }
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);

    std::mutex mu;
    std::condition_variable cv;
    std::deque<Pending> queue;
    std::atomic<uint64_t> served(0), batches(0);

    // one worker per GPU: context created once, buffers kept across batches
    std::vector<std::thread> workers;
    std::atomic<int> ready(0), failed(0);
    for (int dev = 0; dev < so.gpus; ++dev) {
        workers.emplace_back([&, dev] {
            gp_ctx* ctx = nullptr;
            if (gp_create(dev, &ctx) != GP_OK) { fprintf(stderr, "ContigsMerger_b200: GPU %d: %s (no CPU fallback)\n", dev, gp_last_error(nullptr)); ++failed; ++ready; cv.notify_all(); return; }
            ++ready;
            cv.notify_all();
            for (;;) {
                std::vector<Pending> batch;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return g_stop || !queue.empty(); });
                    if (queue.empty() && g_stop) break;
                    // more requests of the same wave are usually on their way (one per Pool worker): wait a moment for them
                    lk.unlock();
                    std::this_thread::sleep_for(std::chrono::milliseconds(so.window_ms));
                    lk.lock();
                    if (queue.empty()) continue;
                    const std::string key = options_key(queue.front().req.opt);
                    for (auto it = queue.begin(); it != queue.end() && (int)batch.size() < so.max_batch;) {
                        if (options_key(it->req.opt) == key) { batch.push_back(std::move(*it)); it = queue.erase(it); } else ++it;
                    }
                }
                std::vector<GapInput> in(batch.size());
                for (size_t k = 0; k < batch.size(); ++k) {
                    const Request& rq = batch[k].req;
                    in[k].fasta_path = (!rq.input.empty() && rq.input[0] != '/' && !rq.cwd.empty()) ? rq.cwd + "/" + rq.input : rq.input;
                }
                std::vector<GapOutput> out;
                std::string err;
                const int rc = merge_gaps(ctx, batch[0].req.opt, in, out, err);
                for (size_t k = 0; k < batch.size(); ++k) {
                    Reply rp;
                    if (rc != GP_OK) { rp.exit_code = 3; rp.err = "ContigsMerger_b200: " + err + "\n"; }
                    else if (!out[k].error.empty()) { rp.exit_code = 3; rp.err = "ContigsMerger_b200: " + out[k].error + "\n"; }
                    else {
                        rp.exit_code = out[k].exit_code; rp.wrote_info = out[k].wrote_info;
                        rp.out = std::move(out[k].stdout_text); rp.info = std::move(out[k].info_text);
                        if (batch[k].req.write_gml) rp.gml = std::move(out[k].gml_text);
                    }
                    send_reply(batch[k].fd, rp);
                    ::close(batch[k].fd);
                }
                served += batch.size();
                ++batches;
            }
            gp_destroy(ctx);
        });
    }
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return ready.load() == so.gpus; });
    }
    if (failed.load() == so.gpus) { g_stop = true; cv.notify_all(); for (auto& t : workers) t.join(); ::unlink(so.socket_path.c_str()); return 3; }
    fprintf(stderr, "ContigsMerger_b200: serving on %s (%d GPU%s, batch window %d ms)\n", so.socket_path.c_str(), so.gpus, so.gpus == 1 ? "" : "s", so.window_ms);

    while (!g_stop) {
        const int fd = ::accept(g_listen_fd, nullptr, nullptr);
        if (fd < 0) { if (g_stop) break; continue; }
        char magic[4];
        uint32_t argc = 0;
        std::vector<std::string> args;
        std::string cwd;
        bool ok = read_all(fd, magic, 4) && !memcmp(magic, "GPM1", 4) && read_all(fd, &argc, 4) && argc < 256;
        for (uint32_t i = 0; ok && i < argc; ++i) { std::string a; ok = get_blob(fd, a, 1u << 16); args.push_back(std::move(a)); }
        ok = ok && get_blob(fd, cwd, 1u << 16);
        if (!ok) { ::close(fd); continue; }
        if (args.size() >= 2 && args[1] == "--shutdown") { Reply rp; send_reply(fd, rp); ::close(fd); g_stop = true; break; }
        std::vector<const char*> av;
        for (const std::string& a : args) av.push_back(a.c_str());
        Pending p;
        p.fd = fd;
        p.req.cwd = cwd;
        if (!parse_request((int)av.size(), av.data(), p.req)) {       // CM/main.cpp:216-220
            Reply rp; rp.exit_code = 1; rp.out = "Wrong input.\n";
            send_reply(fd, rp); ::close(fd);
            continue;
        }
        { std::lock_guard<std::mutex> lk(mu); queue.push_back(std::move(p)); }
        cv.notify_one();
    }
    g_stop = true;
    cv.notify_all();
    for (auto& t : workers) t.join();
    ::close(g_listen_fd);
    ::unlink(so.socket_path.c_str());
    fprintf(stderr, "ContigsMerger_b200: served %llu gaps in %llu batches\n", (unsigned long long)served.load(), (unsigned long long)batches.load());
    return 0;
}

} // namespace gpm
