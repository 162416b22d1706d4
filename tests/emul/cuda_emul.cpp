// cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emul.h).
#include "cuda_emul.h"

#include <sys/mman.h>

namespace emul {

thread_local Ctx ctx;
thread_local unsigned char* dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;
constexpr size_t kDynSmem = 232 * 1024;

struct Fiber {
    ucontext_t uc;
    bool done;
    dim3 tidx;
};
struct Warp {
    int count = 0, gen = 0, nlanes = 0;
    alignas(16) unsigned char buf[32][32];
};
struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int alive = 0, bar_arrived = 0, bar_gen = 0, cur = 0;
    ucontext_t sched;
    const std::function<void()>* body = nullptr;
};
thread_local Block* blk = nullptr;
thread_local std::vector<void*>* stacks = nullptr;

void yield_() {
    Block* b = blk;
    swapcontext(&b->fibers[b->cur].uc, &b->sched);
}

void trampoline() {
    Block* b = blk;
    (*b->body)();
    b = blk;
    b->fibers[b->cur].done = true;
    b->alive--;
    // a thread that exits counts as arrived for any barrier the rest are waiting on
    if (b->alive > 0 && b->bar_arrived >= b->alive) {
        b->bar_arrived = 0;
        b->bar_gen++;
    }
    swapcontext(&b->fibers[b->cur].uc, &b->sched);
}

void* get_stack(size_t i) {
    if (!stacks) stacks = new std::vector<void*>();
    while (stacks->size() <= i) {
        void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::perror("emul mmap"); std::abort(); }
        stacks->push_back(p);
    }
    return (*stacks)[i];
}

void run_block(dim3 grid, dim3 block, unsigned bid, const std::function<void()>& body) {
    Block b;
    blk = &b;
    b.body = &body;
    unsigned n = block.x * block.y * block.z;
    ctx.gdim_ = grid;
    ctx.bdim_ = block;
    ctx.bid_ = dim3(bid % grid.x, (bid / grid.x) % grid.y, bid / (grid.x * grid.y));
    if (!dyn_smem) dyn_smem = static_cast<unsigned char*>(std::aligned_alloc(1024, kDynSmem));
    b.fibers.resize(n);
    b.warps.resize((n + 31) / 32);
    for (unsigned w = 0; w < b.warps.size(); ++w) b.warps[w].nlanes = (int)std::min(32u, n - w * 32);
    b.alive = (int)n;
    for (unsigned i = 0; i < n; ++i) {
        Fiber& f = b.fibers[i];
        f.done = false;
        f.tidx = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = get_stack(i);
        f.uc.uc_stack.ss_size = kStack;
        f.uc.uc_link = nullptr;
        makecontext(&f.uc, trampoline, 0);
    }
    while (b.alive > 0) {
        for (unsigned i = 0; i < n; ++i) {
            if (b.fibers[i].done) continue;
            b.cur = (int)i;
            ctx.tid_ = b.fibers[i].tidx;
            swapcontext(&b.sched, &b.fibers[i].uc);
        }
    }
    blk = nullptr;
}
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    if (smem > kDynSmem) { std::fprintf(stderr, "emul: dynamic smem %zu too large\n", smem); std::abort(); }
    unsigned nblocks = grid.x * grid.y * grid.z;
    if (nblocks == 0 || block.x * block.y * block.z == 0) return;
    static int nthreads_cfg = [] {
        const char* e = std::getenv("MMEGO_EMUL_THREADS");
        int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(v, 32));
    }();
    unsigned nw = std::min<unsigned>(nthreads_cfg, nblocks);
    std::atomic<unsigned> next{0};
    auto worker = [&]() {
        for (;;) {
            unsigned b = next.fetch_add(1);
            if (b >= nblocks) break;
            run_block(grid, block, b, body);
        }
    };
    if (nw <= 1) {
        worker();
    } else {
        std::vector<std::thread> ts;
        for (unsigned i = 0; i < nw; ++i) ts.emplace_back(worker);
        for (auto& t : ts) t.join();
    }
}

void syncthreads() {
    Block* b = blk;
    int gen = b->bar_gen;
    if (++b->bar_arrived >= b->alive) {
        b->bar_arrived = 0;
        b->bar_gen++;
    } else {
        while (blk->bar_gen == gen) yield_();
    }
}

static Warp& my_warp() { return blk->warps[blk->cur / 32]; }
int lane_id() { return blk->cur % 32; }

void warp_barrier() {
    Warp& w = my_warp();
    int gen = w.gen;
    if (++w.count >= w.nlanes) {
        w.count = 0;
        w.gen++;
    } else {
        while (w.gen == gen) yield_();
    }
}

void warp_exchange(const void* in, void* out, size_t bytes, int src_lane) {
    Warp& w = my_warp();
    int lane = lane_id();
    std::memcpy(w.buf[lane], in, bytes);
    warp_barrier();
    if (src_lane < 0 || src_lane >= w.nlanes) src_lane = lane;
    std::memcpy(out, w.buf[src_lane], bytes);
    warp_barrier();
}

void warp_allgather(const void* in, void* out_all, size_t bytes) {
    Warp& w = my_warp();
    std::memcpy(w.buf[lane_id()], in, bytes);
    warp_barrier();
    for (int l = 0; l < 32; ++l) std::memcpy(static_cast<unsigned char*>(out_all) + l * bytes, w.buf[l], bytes);
    warp_barrier();
}

}  // namespace emul
