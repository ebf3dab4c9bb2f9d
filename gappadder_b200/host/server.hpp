// server.hpp -- resident ContigsMerger service: one process holds the CUDA contexts, thin clients with the reference's
// command line send one gap each over a unix socket and get stdout / INFO / tmp.gml bytes back.
//
// Why: GAPPadder launches ContigsMerger once per gap from `Pool(nthreads).map(run_merge, ...)`
// (/root/reference/assemble_gaps.py:296-318, MergeContigs.py:85).  A process that creates its own CUDA context pays
// about a second of start-up for a millisecond of work, and one gap cannot fill a B200.  The server batches the
// requests that arrive together (all of `nthreads` workers' gaps in one launch) and answers each client with exactly
// the bytes the single-gap binary would have produced.  The Python side stays unchanged: `ContigsMerger` on its PATH is
// the client (the same binary; it talks to the server when GAPPADDER_B200_SOCKET names a live socket and runs
// in-process otherwise).
#pragma once
#include <string>
#include <vector>

#include "merger.hpp"

namespace gpm {

struct ServeOptions {
    std::string socket_path;
    int gpus = 1;
    int window_ms = 3;          // after the first request of a batch arrives, wait this long for more
    int max_batch = 4096;       // gaps per merge_gaps call
};

// Parsed command line of one request (shared by the in-process binary, the client and the server).
struct Request {
    MergeOptions opt;
    std::string input;          // as given; the server resolves it against `cwd`
    std::string cwd;
    bool write_gml = true;
};

struct Reply {
    int exit_code = 0;
    bool wrote_info = false;
    std::string out, info, gml, err;
};

// Runs the service until a client sends --shutdown or the process gets SIGTERM / SIGINT.  Returns the exit code.
int serve(const ServeOptions& so);

// Client side: sends argv to the server behind `socket_path`.  Returns false when no server answers (the caller then
// runs in-process); otherwise fills `reply`.
bool request_from_server(const std::string& socket_path, int argc, char** argv, Reply& reply);

// argv -> Request (CM/main.cpp:53-231 flags).  False on a flag the reference rejects ("Wrong input.").
bool parse_request(int argc, const char* const* argv, Request& r);

} // namespace gpm
