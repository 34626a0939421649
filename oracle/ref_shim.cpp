/*
 * ref_shim.cpp — builds the reference's OWN hash/probe runtime into oracle/_ref/libref_groupby.so.
 *
 * TEST INFRASTRUCTURE (see oracle.cpp).  Nothing is copied: QueryEngine/MurmurHash.cpp and
 * QueryEngine/GroupByRuntime.cpp are #included where they lie under $(REF) (= /root/reference) at build time.
 * GroupByRuntime.cpp is written to be included from RuntimeFunctions.cpp (RuntimeFunctions.cpp:2122), which
 * supplies get_empty_key<T>, get_matching_group_value*, dynamic_watchdog(); RuntimeFunctions.cpp itself cannot
 * be compiled here (Boost), so this prelude declares the minimum it needs.  get_matching_group_value is the
 * one-slot match-or-claim step restated from RuntimeFunctions.cpp:1953-2005; the hash (MurmurHash3) and the probe
 * loops (get_group_value, get_group_value_fast, ...) are the reference's own code.
 */
#include <cstdint>
#include <cstring>
#include <limits>

#include "QueryEngine/GpuRtConstants.h"
#include "QueryEngine/BufferCompaction.h"
#include "Shared/funcannotations.h"

template <typename T> inline T get_empty_key();
template <> inline int32_t get_empty_key() { return EMPTY_KEY_32; }
template <> inline int64_t get_empty_key() { return EMPTY_KEY_64; }

template <typename T>
static inline int64_t* get_matching_group_value_t(int64_t* groups_buffer, const uint32_t h, const T* key,
                                                  const uint32_t key_count, const uint32_t row_size_quad) {
  auto off = h * row_size_quad;
  auto row_ptr = reinterpret_cast<T*>(groups_buffer + off);
  if (*row_ptr == get_empty_key<T>()) {
    memcpy(row_ptr, key, key_count * sizeof(T));
    auto row_ptr_i8 = reinterpret_cast<int8_t*>(row_ptr + key_count);
    return reinterpret_cast<int64_t*>(align_to_int64(row_ptr_i8));
  }
  if (memcmp(row_ptr, key, key_count * sizeof(T)) == 0) {
    auto row_ptr_i8 = reinterpret_cast<int8_t*>(row_ptr + key_count);
    return reinterpret_cast<int64_t*>(align_to_int64(row_ptr_i8));
  }
  return nullptr;
}

extern "C" int64_t* get_matching_group_value(int64_t* groups_buffer, const uint32_t h, const int64_t* key,
                                             const uint32_t key_count, const uint32_t key_width,
                                             const uint32_t row_size_quad) {
  switch (key_width) {
    case 4: return get_matching_group_value_t(groups_buffer, h, reinterpret_cast<const int32_t*>(key), key_count, row_size_quad);
    case 8: return get_matching_group_value_t(groups_buffer, h, key, key_count, row_size_quad);
    default: return nullptr;
  }
}
extern "C" int32_t get_matching_group_value_columnar_slot(int64_t*, const uint32_t, const uint32_t, const int64_t*,
                                                          const uint32_t, const uint32_t) { return -1; }
extern "C" int64_t* get_matching_group_value_columnar(int64_t*, const uint32_t, const int64_t*, const uint32_t,
                                                      const size_t) { return nullptr; }
extern "C" bool dynamic_watchdog() { return false; }

#include "QueryEngine/MurmurHash.cpp"
#include "QueryEngine/GroupByRuntime.cpp"

/* The reference's chunk decoders, header-only (QueryEngine/DecodersImpl.h): fixed_width_int_decode,
 * fixed_width_unsigned_decode, fixed_width_small_date_decode (days-encoded DATE), fixed_width_double_decode. */
#include "QueryEngine/DecodersImpl.h"
