// ref_arena.hpp -- a bump allocator for the libraries of oracle/_ref (TEST INFRASTRUCTURE).
//
// Two places of the reference order things by HEAP ADDRESS: DistributeOctTree sorts (key count, ExtractorNode*) pairs
// (src/ORBextractor.cc:654) and peac keeps a node's neighbours in a std::set<PlaneSeg*> (include/peac/AHCPlaneSeg.hpp:212).
// While an arena scope is open every allocation of the library comes from one region that is never reused, so addresses
// grow in allocation order and "address order" becomes "creation order" -- one legal allocator among many, and the
// deterministic one the oracle restates.  operator new/delete are replaced with hidden visibility, i.e. for the
// translation units linked into THIS library only.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <new>

#include <sys/mman.h>

static unsigned char *g_arena = nullptr;
static size_t g_arena_cap = 0, g_arena_top = 0;
static bool g_arena_on = false, g_arena_overflow = false;

static void *arena_alloc(size_t n) {
    if (g_arena_on) {
        size_t at = (g_arena_top + 15) & ~(size_t)15;
        if (at + n <= g_arena_cap) {
            g_arena_top = at + n;
            return g_arena + at;
        }
        g_arena_overflow = true;
    }
    void *p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
static void arena_free(void *p) noexcept {
    if (!p) return;
    if (g_arena && (unsigned char *)p >= g_arena && (unsigned char *)p < g_arena + g_arena_cap) return;  // never reused
    free(p);
}
// hidden: these replace operator new/delete for THIS library only (the templates of the reference are instantiated here)
#define HID __attribute__((visibility("hidden")))
HID void *operator new(size_t n) { return arena_alloc(n); }
HID void *operator new[](size_t n) { return arena_alloc(n); }
HID void operator delete(void *p) noexcept { arena_free(p); }
HID void operator delete[](void *p) noexcept { arena_free(p); }
HID void operator delete(void *p, size_t) noexcept { arena_free(p); }
HID void operator delete[](void *p, size_t) noexcept { arena_free(p); }


// open / close an arena scope; everything allocated inside must be destroyed before ref_arena_end()
static int ref_arena_begin() {
    if (!g_arena) {
        g_arena_cap = (size_t)1 << 31;  // virtual; pages are touched on use
        void *m = mmap(nullptr, g_arena_cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) return -1;
        g_arena = (unsigned char *)m;
    }
    g_arena_top = 0;
    g_arena_overflow = false;
    g_arena_on = true;
    return 0;
}
static int ref_arena_end() {
    g_arena_on = false;
    // give the touched pages back so that a long test session does not accumulate resident memory
    madvise(g_arena, (g_arena_top + 4095) & ~(size_t)4095, MADV_DONTNEED);
    return g_arena_overflow ? -1 : 0;
}
