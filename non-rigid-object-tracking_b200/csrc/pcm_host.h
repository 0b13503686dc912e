// pcm_host.h -- host-side helpers of libpcm_b200.so: a small persistent thread pool used to
// stage HOST buffers (crop rows -> pinned memory, strided mask scatter/gather) in parallel,
// so that the host<->device copies of the e2e path run at PCIe speed instead of at the speed
// of one memcpy thread.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pcm {

class HostPool {
public:
    static HostPool& instance() {
        static HostPool pool;
        return pool;
    }
    int size() const { return (int)workers_.size() + 1; }

    // Calls fn(i) for i in [0, n) on the pool's threads plus the caller; returns when all are done.
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        std::unique_lock<std::mutex> guard(submit_);   // one job at a time
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = n;
            ++generation_;
        }
        cv_.notify_all();
        run_items();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        int n = 0;
        if (const char* e = getenv("PCM_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = (int)(hw / 2);
            if (n > 8) n = 8;
        }
        if (n < 1) n = 1;
        for (int i = 1; i < n; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run_items() {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void worker() {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            run_items();
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_, submit_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, pending_ = 0;
    std::atomic<int> next_{0};
    unsigned long generation_ = 0;
    bool stop_ = false;
};

}  // namespace pcm
