// tests/gate_test.cpp -- DeviceGate (gappadder_b200/host/device_gate.hpp) under several "mergers": the rules it exists for
// hold at every moment -- one pairwise phase at a time, one relax launch at a time, and no pairwise phase begins while a
// merger that finished its pairwise phase has not enqueued its relax kernel yet -- and nobody deadlocks, whichever of the
// three exits a merger takes after its pairwise phase (relax launch, cancelled relax, launch that fails before the hook).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>
#include <vector>

#include "../gappadder_b200/host/device_gate.hpp"

int main()
{
    gpm::DeviceGate gate;
    std::atomic<int> in_pairwise(0), in_relax(0), pending(0), violations(0), done(0);
    auto nap = [](int us) { std::this_thread::sleep_for(std::chrono::microseconds(us)); };
    auto merger = [&](int id) {
        std::mt19937 rng(1234u + (unsigned)id);
        for (int chunk = 0; chunk < 60; ++chunk) {
            nap((int)(rng() % 300));                                  // host phase: nodes, pair list
            gate.begin_pairwise();
            if (in_pairwise.fetch_add(1) != 0) ++violations;         // one pairwise phase at a time
            if (pending.load() != 0) ++violations;                   // never while a relax launch is announced
            nap((int)(rng() % 400));
            in_pairwise.fetch_sub(1);
            const int exit_kind = (int)(rng() % 4);                  // 0,1: relax launch; 2: cancelled; 3: no relax announced
            if (exit_kind == 3) { gate.end_pairwise(false); continue; }
            pending.fetch_add(1);                                     // before the gate opens: others must see it
            gate.end_pairwise(true);
            nap((int)(rng() % 200));                                  // graph phase
            if (exit_kind == 2) { pending.fetch_sub(1); gate.cancel_relax(); continue; }
            gate.begin_relax();
            if (in_relax.fetch_add(1) != 0) ++violations;            // one relax launch at a time
            nap((int)(rng() % 50));                                   // until the kernel is enqueued
            pending.fetch_sub(1);
            gate.relax_launched();                                    // (exit 1: "the call failed before launching" is the same call)
            nap((int)(rng() % 400));                                  // the launch runs; pairwise phases of others may start now
            in_relax.fetch_sub(1);
            gate.end_relax();
        }
        ++done;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < 4; ++t) th.emplace_back(merger, t);
    for (auto& t : th) t.join();
    std::printf("mergers done %d violations %d\n", done.load(), violations.load());
    return done.load() == 4 && violations.load() == 0 ? 0 : 1;
}
