// host_pool.hpp — a handful of persistent host threads for the per-batch host work (building the K plans of a
// batch): spawning threads per call costs as much as the work they do (a C2 plan is ~15 us).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sgufp {

class HostPool {
public:
    explicit HostPool(int workers) {
        for (int t = 0; t < workers; t++) threads_.emplace_back([this, t] { loop(t + 1); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }   // the caller works too
    // fn(part, parts) runs once per part on `parts` threads (the caller is part 0); returns when all are done
    void run(int parts, const std::function<void(int, int)> &fn) {
        parts = std::max(1, std::min(parts, size()));
        if (parts == 1) { fn(0, 1); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; parts_ = parts; pending_ = parts - 1; gen_++;
        }
        cv_.notify_all();
        fn(0, parts);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn = nullptr;
            int parts = 0;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (id >= parts_) continue;              // not needed for this job
                fn = fn_; parts = parts_;
            }
            (*fn)(id, parts);
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int parts_ = 0, pending_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

}  // namespace sgufp
