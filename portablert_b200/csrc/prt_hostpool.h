// prt_hostpool.h -- a small pool of host threads for the staging copies of the host entry points.
//
// The reference API hands over pageable memory (std::vector<Ray> in, std::vector<HitReg> out,
// backend.hpp:75-83).  DMA needs page-locked memory, so such buffers are staged through a pinned
// ring; one core copies ~10 GB/s, a PCIe 5 x16 link moves ~55 GB/s, hence several threads.  The
// pool is created once per context (no thread creation on the call path) and runs one parallel
// loop at a time.  (Measured on the B200 host, 30 M rays in, 30 M 16-byte records out: 104 / 74 /
// 39 / 33 / 32 ms with 1 / 2 / 4 / 8 / 16 copy threads -- bound by the host's memory bandwidth from
// 8 threads on; hand-written MOVNTDQ copies were 1.4-2.5x SLOWER than glibc's memcpy there.)
#pragma once

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace prt {

class HostPool {
  public:
	explicit HostPool(int workers) {
		for (int k = 0; k < workers; ++k)
			th_.emplace_back([this, k] { loop(k + 1); });
	}
	~HostPool() {
		{
			std::lock_guard<std::mutex> l(m_);
			stop_ = true;
		}
		cv_.notify_all();
		for (auto &t : th_)
			t.join();
	}
	int width() const { return (int)th_.size() + 1; }

	// fn(part, parts) for part in [0, parts), parts = width(): the calling thread takes part 0
	void run(const std::function<void(int, int)> &fn) {
		const int parts = width();
		if (parts == 1) {
			fn(0, 1);
			return;
		}
		{
			std::lock_guard<std::mutex> l(m_);
			fn_ = &fn;
			pending_ = parts - 1;
			++epoch_;
		}
		cv_.notify_all();
		fn(0, parts);
		std::unique_lock<std::mutex> l(m_);
		done_.wait(l, [&] { return pending_ == 0; });
		fn_ = nullptr;
	}

	// memcpy split into cache-line aligned parts
	void copy(void *dst, const void *src, size_t bytes) {
		if (bytes < (1u << 20) || width() == 1) {
			std::memcpy(dst, src, bytes);
			return;
		}
		run([&](int part, int parts) {
			const size_t per = ((bytes + parts - 1) / parts + 63) & ~size_t(63);
			const size_t off = (size_t)part * per;
			if (off < bytes)
				std::memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off,
				            std::min(per, bytes - off));
		});
	}

  private:
	void loop(int part) {
		uint64_t seen = 0;
		for (;;) {
			const std::function<void(int, int)> *fn = nullptr;
			{
				std::unique_lock<std::mutex> l(m_);
				cv_.wait(l, [&] { return stop_ || epoch_ != seen; });
				if (stop_)
					return;
				seen = epoch_;
				fn = fn_;
			}
			(*fn)(part, width());
			{
				std::lock_guard<std::mutex> l(m_);
				--pending_;
			}
			done_.notify_one();
		}
	}
	std::vector<std::thread> th_;
	std::mutex m_;
	std::condition_variable cv_, done_;
	const std::function<void(int, int)> *fn_ = nullptr;
	uint64_t epoch_ = 0;
	int pending_ = 0;
	bool stop_ = false;
};

} // namespace prt
