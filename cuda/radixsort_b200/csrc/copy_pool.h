// copy_pool.h -- a few host threads that split one large memcpy between them.  Used to move
// pageable host arrays through pinned staging buffers faster than one core can copy.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace b200sort {

class CopyPool {
public:
    void copy(void *dst, const void *src, size_t bytes) {
        ensure_started();
        const int parts = (int)workers_.size() + 1;
        const size_t per = (((bytes + parts - 1) / parts) + 4095) & ~size_t(4095);  // >= 4096 for bytes > 0
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = static_cast<char *>(dst);
            src_ = static_cast<const char *>(src);
            bytes_ = bytes;
            per_ = per;
            pending_ = (int)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        run_part(0);  // the caller takes the first slice
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }

private:
    void ensure_started() {
        if (started_) return;
        started_ = true;
        unsigned hw = std::thread::hardware_concurrency();
        int n = (int)std::min<unsigned>(hw > 2 ? hw / 2 : 1, 6) - 1;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { worker(i + 1); });
    }
    void run_part(int part) {
        const size_t off = per_ * (size_t)part;
        if (off < bytes_) memcpy(dst_ + off, src_ + off, std::min(per_, bytes_ - off));
    }
    void worker(int part) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            run_part(part);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --pending_;
            }
            done_cv_.notify_one();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> workers_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false, started_ = false;
};

}  // namespace b200sort
