// pcm_host.h -- host-side helpers of libpcm_b200.so: a small persistent thread pool used to
// stage HOST buffers (crop rows -> pinned memory, strided mask scatter/gather) in parallel,
// so that the host<->device copies of the e2e path run at PCIe speed instead of at the speed
// of one memcpy thread.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <memory>
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
    // Every call publishes its own Job object, so a worker that wakes up late only ever sees a
    // finished job (next >= n) or the current one -- never a mix of two calls' fields.
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        std::unique_lock<std::mutex> guard(submit_);   // one job at a time
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        job->n = n;
        bool wake;
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = job;
            generation_.fetch_add(1, std::memory_order_release);
            wake = sleepers_ > 0;
        }
        if (wake) cv_.notify_all();
        run_items(*job);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return job->done.load() == n; });
    }

private:
    HostPool() {
        int n = 0;
        if (const char* e = getenv("PCM_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            // the strided mask scatter / gather of a 1080p frame is bound by per-core memory bandwidth on the B200 hosts
            // (16 cores, ~10 GB/s each: measured 0.14 ms with 6 threads, less with more), so take three quarters of the
            // cores this rank may use, up to 12.  Under a one-process-per-GPU launcher the cores are shared by
            // LOCAL_WORLD_SIZE ranks.
            const unsigned hw = std::thread::hardware_concurrency();
            int ranks = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e);
            if (ranks < 1) ranks = 1;
            const int cores = (int)hw / ranks;
            n = cores <= 6 ? cores : cores * 3 / 4;      // the calling thread works too: a rank with four cores uses all four
            if (n > 12) n = 12;
            if (cores < 4) spin_us_ = 50;     // no spare core to spin on
        }
        if (const char* e = getenv("PCM_HOST_SPIN_US")) spin_us_ = atoi(e);
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
    struct Job {
        const std::function<void(int)>* fn = nullptr;
        int n = 0;
        std::atomic<int> next{0};
        std::atomic<int> done{0};
    };
    void run_items(Job& job) {
        for (;;) {
            const int i = job.next.fetch_add(1);
            if (i >= job.n) break;
            (*job.fn)(i);
            if (job.done.fetch_add(1) + 1 == job.n) {
                std::lock_guard<std::mutex> lk(m_);   // pairs with the waiter's predicate check
                done_.notify_all();
            }
        }
    }
    // A worker that has just finished a job spins for a short while before it sleeps on the
    // condition variable: the host entry points arrive in bursts (update, IoU, update, ...) a few
    // hundred microseconds apart, and a futex wake-up costs more than the staging work it starts.
    void worker() {
        unsigned long seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            bool hot = false;
            for (int spin = 0; !hot; ++spin) {
                if (generation_.load(std::memory_order_acquire) != seen) { hot = true; break; }
                if ((spin & 63) == 63 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
                cpu_relax();
            }
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                if (!hot) {
                    ++sleepers_;
                    cv_.wait(lk, [&] { return stop_ || generation_.load() != seen; });
                    --sleepers_;
                }
                if (stop_) return;
                seen = generation_.load();
                job = job_;
            }
            if (job) run_items(*job);
        }
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    std::vector<std::thread> workers_;
    std::mutex m_, submit_;
    std::condition_variable cv_, done_;
    std::shared_ptr<Job> job_;
    std::atomic<unsigned long> generation_{0};
    int sleepers_ = 0;
    int spin_us_ = 400;
    bool stop_ = false;
};

}  // namespace pcm
