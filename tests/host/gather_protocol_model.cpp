// CPU model of the multi-GPU en-face gather protocol (octproz_b200/csrc/oct_device.cuh GatherDev, k_fused.cuh, k_aux.cu
// enface_consume_kernel).  TEST-ONLY: it restates the SYNCHRONISATION of the device code with std::atomic
// release / acquire in place of the system-scope PTX, so that the protocol's claims can be checked without a GPU and at a world size
// the GPU suite does not reach:
//   * no dead-lock with every rank's kernels in stream order (main(s), consume(s), main(s+1), ...), the prologue of main(s+1)
//     (its acknowledgement wait) overlapping consume(s) as under programmatic dependent launch;
//   * a consumer only ever copies a COMPLETE frame of ITS sequence number, however far the ranks drift apart (three frame buffers,
//     producers wait for ack >= seq - 3);
//   * the frame data itself is plain memory: ThreadSanitizer sees a happens-before chain
//     stores of main(s) -> arrived[] (release) -> consume(s) reads (acquire) -> ack[] (release) -> main(s+3) overwrites (acquire).
// argv: world steps [mode]
//   inorder       (default) the shipped form: consume(s) sits between main(s) and main(s+1) on the compute stream.  Stream order alone
//                 then keeps all ranks within one step of each other (main(s+1) needs the own consume(s), which needs every rank's
//                 main(s)); the acknowledgement wait is a second line of defence that never blocks.
//   async         the consumer kernels on a stream of their own (measured and dropped on the GPU because a saturated compute stream can
//                 starve them): here the acknowledgements are what keeps a fast producer from overwriting a frame
//                 a slow consumer still reads -- the model must stay clean.
//   async-noflow  the same without the acknowledgement wait: negative control, torn frames MUST be detected (exit code 2).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace {

constexpr int kFrames = 3;        // OCT_GATHER_FRAMES
constexpr int kSlab = 64;         // en-face values per rank and frame (A * B_local)

struct Window {                   // one per rank: [arrived[producer]] [ack[consumer]] [frame 0..2]
	std::vector<std::atomic<uint32_t>> arrived, ack;
	std::vector<uint32_t> frame[kFrames];          // plain memory on purpose
	explicit Window(int world) : arrived(world), ack(world) {
		for (auto& a : arrived) a.store(0);
		for (auto& a : ack) a.store(0);
		for (auto& f : frame) f.assign((size_t)world * kSlab, 0u);
	}
};

struct Rank {                     // stream order inside one rank
	std::atomic<uint32_t> produced{0}, consumed{0};
};

std::atomic<bool> g_abort{false};
std::atomic<long> g_torn{0}, g_timeouts{0};

bool spin_ge(const std::atomic<uint32_t>& w, uint32_t want) {
	const auto t0 = std::chrono::steady_clock::now();
	while ((int32_t)(w.load(std::memory_order_acquire) - want) < 0) {
		if (g_abort.load(std::memory_order_relaxed)) return false;
		if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) { g_timeouts++; g_abort = true; return false; }
		std::this_thread::yield();
	}
	return true;
}

void dawdle(std::mt19937& rng, int slowness) {
	const int k = (int)(rng() % 8u);
	if (k < slowness) std::this_thread::sleep_for(std::chrono::microseconds(20 + rng() % 200));
	else if (k == 7) std::this_thread::yield();
}

}  // namespace

int main(int argc, char** argv) {
	const int world = argc > 1 ? std::atoi(argv[1]) : 8;
	const uint32_t steps = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 2000u;
	const char* mode = argc > 3 ? argv[3] : "inorder";
	const bool flow = std::strcmp(mode, "async-noflow") != 0;
	const bool inorder = std::strcmp(mode, "inorder") == 0;
	std::vector<Window*> win;
	std::vector<Rank*> rk;
	for (int r = 0; r < world; ++r) { win.push_back(new Window(world)); rk.push_back(new Rank()); }

	std::vector<std::thread> th;
	for (int r = 0; r < world; ++r) {
		// ---- the main kernels of rank r: prologue = acknowledgement wait, body = peer stores spread over the kernel, tail = publish
		th.emplace_back([&, r] {
			std::mt19937 rng(1000u + (unsigned)r);
			const int slowness = (r == world - 1) ? 5 : (r == 0 ? 0 : 2);          // rank 0 runs ahead, the last rank dawdles
			for (uint32_t s = 1; s <= steps && !g_abort; ++s) {
				if (flow && s > (uint32_t)kFrames)                                  // gather_wait_acks: lane c looks at consumer c, own window
					for (int c = 0; c < world; ++c) if (!spin_ge(win[r]->ack[c], s - kFrames)) return;
				if (inorder && !spin_ge(rk[r]->consumed, s - 1)) return;            // griddepcontrol.wait: consume(s-1) of this rank is complete
				for (int blk = 0; blk < kSlab; blk += 8) {                          // one block of 8 lines at a time, to every rank
					for (int d = 0; d < world; ++d)
						for (int i = 0; i < 8; ++i) win[d]->frame[s % kFrames][(size_t)r * kSlab + blk + i] = (s << 8) | (uint32_t)r;
					if ((blk & 24) == 0) dawdle(rng, slowness);
				}
				for (int d = 0; d < world; ++d) win[d]->arrived[r].store(s, std::memory_order_release);     // gather_publish (last CTA)
				rk[r]->produced.store(s, std::memory_order_release);
			}
		});
		// ---- the consumer kernels of rank r, each in stream order behind the main kernel of the same sequence number
		th.emplace_back([&, r] {
			std::mt19937 rng(2000u + (unsigned)r);
			const int slowness = (r == world - 1) ? 5 : 1;
			std::vector<uint32_t> display((size_t)world * kSlab);
			for (uint32_t s = 1; s <= steps && !g_abort; ++s) {
				if (!spin_ge(rk[r]->produced, s)) return;                           // stream order (dependent launch: body after main(s))
				for (int p = 0; p < world; ++p) if (!spin_ge(win[r]->arrived[p], s)) return;
				dawdle(rng, slowness);
				std::memcpy(display.data(), win[r]->frame[s % kFrames].data(), display.size() * sizeof(uint32_t));
				for (int p = 0; p < world; ++p)
					for (int i = 0; i < kSlab; ++i)
						if (display[(size_t)p * kSlab + i] != ((s << 8) | (uint32_t)p)) { g_torn++; break; }
				for (int p = 0; p < world; ++p) win[p]->ack[r].store(s, std::memory_order_release);        // ack[consumer] in every producer's window
				rk[r]->consumed.store(s, std::memory_order_release);
			}
		});
	}
	for (auto& t : th) t.join();
	const long torn = g_torn.load(), to = g_timeouts.load();
	std::printf("world %d steps %u mode %s torn_frames %ld timeouts %ld\n", world, steps, mode, torn, to);
	for (auto* w : win) delete w;
	for (auto* k : rk) delete k;
	if (to) return 3;
	return torn ? 2 : 0;
}
