/* oracle/zero_new.cc — zero-initialising global operator new/delete (TEST INFRASTRUCTURE ONLY).
 * Linked into every oracle/_ref binary: the reference reads memory it never wrote
 * (adjMatrix.cc:36-43 leaves row_offset[0] unset; serial/lib/multiplyOut.cc:27-33 accumulates onto fresh new[]
 * buffers). calloc makes those reads see zeros, which is what the author's runs saw on a virgin heap. */
#include <cstdlib>
#include <new>
void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
