// device_gate.hpp -- who may use one GPU when several host threads (mergers, one context each) drive it.
//
// Two classes of device phases: PAIRWISE (table upload, quick check, the pairwise launch: persistent kernels that fill
// every SM until their work queue is empty) and RELAX (gp_relax_chains: its bulk fills the chip, its tail is a few long
// chains -- a critical path of some 25 ms whatever the batch size -- on a handful of CTAs while idle CTAs exit).
// Rules:
//   * one pairwise phase at a time, one relax launch at a time;
//   * a merger that finished its pairwise phase and is about to launch its relax chains (it only has to build the
//     graphs first) holds back every other merger's pairwise phase until the relax kernel is ENQUEUED
//     (gp_set_relax_launch_hook): the relax CTAs take the SMs first, the next chunk's pairwise CTAs fill them as
//     relax CTAs exit, and the relax tail runs beside the pairwise kernel instead of an idle chip.
// Nothing here affects results: gaps are independent and every launch is what it would be alone.
#pragma once
#include <condition_variable>
#include <mutex>

namespace gpm {

class DeviceGate {
public:
    void begin_pairwise()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !pairwise_busy_ && relax_pending_ == 0; });
        pairwise_busy_ = true;
    }
    // relax_follows: this merger will call begin_relax() shortly (or cancel_relax() if it has no chain after all)
    void end_pairwise(bool relax_follows)
    {
        std::lock_guard<std::mutex> lk(mu_);
        pairwise_busy_ = false;
        if (relax_follows) ++relax_pending_;
        cv_.notify_all();
    }
    void begin_relax()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !relax_busy_; });
        relax_busy_ = true;
    }
    void relax_launched()                       // the kernel is in its stream: pairwise phases may queue behind it
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (relax_pending_ > 0) --relax_pending_;
        cv_.notify_all();
    }
    void cancel_relax() { relax_launched(); }   // announced by end_pairwise(true), not needed after all
    void end_relax()
    {
        std::lock_guard<std::mutex> lk(mu_);
        relax_busy_ = false;
        cv_.notify_all();
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    bool pairwise_busy_ = false, relax_busy_ = false;
    int relax_pending_ = 0;
};

} // namespace gpm
