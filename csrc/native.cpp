// Host-side C++ runtime for bigslice_b200 (pybind11 module `_bsnative`).
//
//  * EvalCore — the evaluator's dependency bookkeeping (phase-aware waitlist:
//    consumers wait on whole producer groups).  Python drives task state and
//    executor calls; this core decides what becomes runnable.  Mirrors
//    bigslice_b200/exec/eval.py:_PyCore; role of exec/eval.go:272-451 `state`.
//  * crc32        — slice-by-8 CRC32 (IEEE) for the frame codec.
//  * murmur3_*    — seeded MurmurHash3_x86_32 over strings / integer buffers
//    (the host half of the bit-exact partitioner).
//  * merge_bound  — helper for batched k-way merges.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace py = pybind11;

// ---------------------------------------------------------------------------
class EvalCore {
 public:
  void set_count(int64_t task, int n) {
    counts_[task] = n;
    if (n == 0) todo_.push_back(task);
  }
  bool has_group(int64_t g) const { return group_missing_.count(g) != 0; }
  void add_group(int64_t g, const std::vector<int64_t>& missing) {
    auto& s = group_missing_[g];
    for (int64_t t : missing) {
      s.insert(t);
      task_groups_[t].push_back(g);
    }
  }
  void wait_on(int64_t task, int64_t g) { group_waiters_[g].push_back(task); }
  void schedule(int64_t task) { todo_.push_back(task); }
  void done_ok(int64_t task) {
    auto it = task_groups_.find(task);
    if (it == task_groups_.end()) return;
    std::vector<int64_t> groups = std::move(it->second);
    task_groups_.erase(it);
    for (int64_t g : groups) {
      auto gm = group_missing_.find(g);
      if (gm == group_missing_.end()) continue;
      gm->second.erase(task);
      if (!gm->second.empty()) continue;
      group_missing_.erase(gm);
      auto gw = group_waiters_.find(g);
      if (gw == group_waiters_.end()) continue;
      std::vector<int64_t> waiters = std::move(gw->second);
      group_waiters_.erase(gw);
      for (int64_t c : waiters) {
        int& n = counts_[c];
        if (--n == 0) todo_.push_back(c);
      }
    }
  }
  std::vector<int64_t> take_todo() {
    std::vector<int64_t> t;
    t.swap(todo_);
    return t;
  }
  size_t num_waiting_groups() const { return group_missing_.size(); }

 private:
  std::unordered_map<int64_t, int> counts_;
  std::unordered_map<int64_t, std::unordered_set<int64_t>> group_missing_;
  std::unordered_map<int64_t, std::vector<int64_t>> group_waiters_;
  std::unordered_map<int64_t, std::vector<int64_t>> task_groups_;
  std::vector<int64_t> todo_;
};

// ---------------------------------------------------------------------------
static uint32_t crc_tab[8][256];
static bool crc_ready = false;
static void crc_init() {
  if (crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xff];
  crc_ready = true;
}
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  crc_init();
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    crc = crc_tab[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    uint32_t lo = (uint32_t)v ^ crc, hi = (uint32_t)(v >> 32);
    crc = crc_tab[7][lo & 0xff] ^ crc_tab[6][(lo >> 8) & 0xff] ^ crc_tab[5][(lo >> 16) & 0xff] ^
          crc_tab[4][lo >> 24] ^ crc_tab[3][hi & 0xff] ^ crc_tab[2][(hi >> 8) & 0xff] ^
          crc_tab[1][(hi >> 16) & 0xff] ^ crc_tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = crc_tab[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return ~crc;
}
static uint32_t py_crc32(py::buffer b, uint32_t value) {
  py::buffer_info info = b.request();
  size_t n = (size_t)info.size * (size_t)info.itemsize;
  return crc32_update(value, static_cast<const uint8_t*>(info.ptr), n);
}

// GF(2) crc32_combine (zlib algorithm) for per-block device CRCs.
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  while (vec) {
    if (vec & 1) sum ^= *mat;
    vec >>= 1;
    ++mat;
  }
  return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat) {
  for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]);
}
static uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;
  uint32_t row = 1;
  for (int n = 1; n < 32; ++n) {
    odd[n] = row;
    row <<= 1;
  }
  gf2_square(even, odd);
  gf2_square(odd, even);
  do {
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}
static uint32_t py_crc32_combine_blocks(py::array_t<uint32_t> crcs, uint64_t total, uint64_t block) {
  auto r = crcs.unchecked<1>();
  uint32_t crc = 0;
  uint64_t left = total;
  for (py::ssize_t i = 0; i < r.shape(0); ++i) {
    uint64_t len = left < block ? left : block;
    crc = (i == 0) ? r(i) : crc32_combine(crc, r(i), len);
    left -= len;
  }
  return crc;
}

// ---------------------------------------------------------------------------
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t murmur3_32(const uint8_t* data, size_t len, uint32_t seed) {
  uint32_t h = seed;
  const size_t nblocks = len / 4;
  for (size_t i = 0; i < nblocks; ++i) {
    uint32_t k;
    memcpy(&k, data + 4 * i, 4);
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  const uint8_t* tail = data + 4 * nblocks;
  uint32_t k = 0;
  switch (len & 3) {
    case 3: k ^= (uint32_t)tail[2] << 16; [[fallthrough]];
    case 2: k ^= (uint32_t)tail[1] << 8; [[fallthrough]];
    case 1:
      k ^= tail[0];
      k *= 0xcc9e2d51u;
      k = rotl32(k, 15);
      k *= 0x1b873593u;
      h ^= k;
  }
  h ^= (uint32_t)len;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
static py::array_t<uint32_t> py_murmur3_strings(const std::vector<py::bytes>& items, uint32_t seed) {
  py::array_t<uint32_t> out((py::ssize_t)items.size());
  auto w = out.mutable_unchecked<1>();
  for (size_t i = 0; i < items.size(); ++i) {
    char* p;
    py::ssize_t n;
    PyBytes_AsStringAndSize(items[i].ptr(), &p, &n);
    w((py::ssize_t)i) = murmur3_32(reinterpret_cast<const uint8_t*>(p), (size_t)n, seed);
  }
  return out;
}
static py::array_t<uint32_t> py_murmur3_fixed(py::buffer b, int width, uint32_t seed) {
  py::buffer_info info = b.request();
  size_t nbytes = (size_t)info.size * (size_t)info.itemsize;
  size_t n = nbytes / (size_t)width;
  py::array_t<uint32_t> out((py::ssize_t)n);
  auto w = out.mutable_unchecked<1>();
  const uint8_t* p = static_cast<const uint8_t*>(info.ptr);
  for (size_t i = 0; i < n; ++i) w((py::ssize_t)i) = murmur3_32(p + i * width, (size_t)width, seed);
  return out;
}
static uint32_t py_murmur3_bytes(py::bytes b, uint32_t seed) {
  char* p;
  py::ssize_t n;
  PyBytes_AsStringAndSize(b.ptr(), &p, &n);
  return murmur3_32(reinterpret_cast<const uint8_t*>(p), (size_t)n, seed);
}

// ---------------------------------------------------------------------------
// Host data-path kernels for the CPU executors (called with the GIL released, so
// the evaluator's thread pool scales across cores): murmur3 of fixed-width keys,
// stable partition permutation, hash aggregation of <int64 key, numeric value>.
// Tensors are passed as raw addresses (torch `data_ptr()`); the Python side
// guarantees contiguity, dtype and lifetime.
static inline uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
static inline uint32_t mm_round(uint32_t h, uint32_t k) {
  k *= 0xcc9e2d51u;
  k = rotl32(k, 15);
  k *= 0x1b873593u;
  h ^= k;
  h = rotl32(h, 13);
  return h * 5 + 0xe6546b64u;
}
static inline uint32_t murmur3_u64(uint64_t v, uint32_t seed) {
  uint32_t h = mm_round(seed, (uint32_t)v);
  h = mm_round(h, (uint32_t)(v >> 32));
  return fmix32(h ^ 8u);
}
static inline uint32_t murmur3_u32(uint32_t v, uint32_t seed) { return fmix32(mm_round(seed, v) ^ 4u); }

// out[i] = murmur3 of the 8-byte (width 8) or low 4-byte (width 4) value in[i]; in/out int64.
static void host_hash_i64(uintptr_t in_, uintptr_t out_, int64_t n, int width, uint32_t seed, bool xor_into) {
  const int64_t* in = reinterpret_cast<const int64_t*>(in_);
  int64_t* out = reinterpret_cast<int64_t*>(out_);
  if (width == 8) {
    for (int64_t i = 0; i < n; ++i) {
      uint32_t h = murmur3_u64((uint64_t)in[i], seed);
      out[i] = xor_into ? (out[i] ^ (int64_t)h) : (int64_t)h;
    }
  } else {
    for (int64_t i = 0; i < n; ++i) {
      uint32_t h = murmur3_u32((uint32_t)in[i], seed);
      out[i] = xor_into ? (out[i] ^ (int64_t)h) : (int64_t)h;
    }
  }
}

// Stable counting sort of row indices by destination id: perm groups the rows of
// partition 0, then 1, ...; counts[p] = rows of partition p.  Returns -1 on an
// id outside [0, npart).
static int64_t host_partition_perm(uintptr_t ids_, int64_t n, int64_t npart, uintptr_t perm_, uintptr_t counts_) {
  const int64_t* ids = reinterpret_cast<const int64_t*>(ids_);
  int64_t* perm = reinterpret_cast<int64_t*>(perm_);
  int64_t* counts = reinterpret_cast<int64_t*>(counts_);
  for (int64_t p = 0; p < npart; ++p) counts[p] = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t d = ids[i];
    if (d < 0 || d >= npart) return -1;
    ++counts[d];
  }
  std::vector<int64_t> cur((size_t)npart);
  int64_t run = 0;
  for (int64_t p = 0; p < npart; ++p) {
    cur[(size_t)p] = run;
    run += counts[p];
  }
  for (int64_t i = 0; i < n; ++i) perm[cur[(size_t)ids[i]]++] = i;
  return n;
}

template <typename V>
struct HostCombine {
  static inline V add(V a, V b) {
    if constexpr (std::is_integral<V>::value) {  // wraps like the device atomics (no signed-overflow UB)
      using U = typename std::make_unsigned<V>::type;
      return (V)((U)a + (U)b);
    } else {
      return a + b;
    }
  }
  static inline V apply(int op, V a, V b) { return op == 0 ? add(a, b) : (op == 1 ? (b < a ? b : a) : (b > a ? b : a)); }
};

// Open-addressing aggregation; writes the distinct keys and combined values
// (unordered) to out_keys/out_vals (capacity n) and returns their number.
template <typename V>
static int64_t host_combine_t(const int64_t* keys, const V* vals, int64_t n, int op, int64_t* out_keys, V* out_vals) {
  if (n <= 0) return 0;
  size_t cap = 1024;
  while (cap < (size_t)n / 4) cap <<= 1;  // grows on demand: most inputs have far fewer keys than rows
  std::vector<int64_t> tk(cap);
  std::vector<V> tv(cap);
  std::vector<uint8_t> used(cap, 0);
  size_t count = 0;
  auto mix = [](uint64_t x) -> uint64_t {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
  };
  auto grow = [&]() {
    size_t ncap = cap * 2;
    std::vector<int64_t> nk(ncap);
    std::vector<V> nv(ncap);
    std::vector<uint8_t> nu(ncap, 0);
    for (size_t i = 0; i < cap; ++i) {
      if (!used[i]) continue;
      size_t s = (size_t)mix((uint64_t)tk[i]) & (ncap - 1);
      while (nu[s]) s = (s + 1) & (ncap - 1);
      nu[s] = 1;
      nk[s] = tk[i];
      nv[s] = tv[i];
    }
    tk.swap(nk);
    tv.swap(nv);
    used.swap(nu);
    cap = ncap;
  };
  for (int64_t i = 0; i < n; ++i) {
    const int64_t k = keys[i];
    size_t s = (size_t)mix((uint64_t)k) & (cap - 1);
    while (true) {
      if (!used[s]) {
        used[s] = 1;
        tk[s] = k;
        tv[s] = vals[i];
        if (++count * 2 > cap) grow();
        break;
      }
      if (tk[s] == k) {
        tv[s] = HostCombine<V>::apply(op, tv[s], vals[i]);
        break;
      }
      s = (s + 1) & (cap - 1);
    }
  }
  int64_t m = 0;
  for (size_t i = 0; i < cap; ++i) {
    if (used[i]) {
      out_keys[m] = tk[i];
      out_vals[m] = tv[i];
      ++m;
    }
  }
  return m;
}

// vtype: 0 int32, 1 int64, 2 float32, 3 float64; op: 0 add, 1 min, 2 max.
static int64_t host_hash_combine(uintptr_t keys, uintptr_t vals, int64_t n, int vtype, int op,
                                 uintptr_t out_keys, uintptr_t out_vals) {
  const int64_t* k = reinterpret_cast<const int64_t*>(keys);
  int64_t* ok = reinterpret_cast<int64_t*>(out_keys);
  switch (vtype) {
    case 0: return host_combine_t<int32_t>(k, reinterpret_cast<const int32_t*>(vals), n, op, ok, reinterpret_cast<int32_t*>(out_vals));
    case 1: return host_combine_t<int64_t>(k, reinterpret_cast<const int64_t*>(vals), n, op, ok, reinterpret_cast<int64_t*>(out_vals));
    case 2: return host_combine_t<float>(k, reinterpret_cast<const float*>(vals), n, op, ok, reinterpret_cast<float*>(out_vals));
    case 3: return host_combine_t<double>(k, reinterpret_cast<const double*>(vals), n, op, ok, reinterpret_cast<double*>(out_vals));
  }
  return -1;
}

// ---------------------------------------------------------------------------
// Object (str / bytes) key columns on the host: hashing and aggregation without
// a Python-level loop.  A str's UTF-8 form is cached inside the object, so views
// of it stay valid while the sequence is alive.
static inline bool object_bytes(PyObject* o, const char** p, Py_ssize_t* l) {
  if (o == Py_None) {
    *p = "";
    *l = 0;
    return true;
  }
  if (PyUnicode_Check(o)) {
    *p = PyUnicode_AsUTF8AndSize(o, l);
    if (*p == nullptr) throw py::error_already_set();
    return true;
  }
  if (PyBytes_Check(o)) {
    char* q;
    if (PyBytes_AsStringAndSize(o, &q, l) != 0) throw py::error_already_set();
    *p = q;
    return true;
  }
  if (PyByteArray_Check(o)) {
    *p = PyByteArray_AS_STRING(o);
    *l = PyByteArray_GET_SIZE(o);
    return true;
  }
  return false;
}

static py::array_t<int64_t> py_murmur3_objects(py::sequence items, uint32_t seed) {
  const py::ssize_t n = py::len(items);
  py::array_t<int64_t> out(n);
  auto w = out.mutable_unchecked<1>();
  for (py::ssize_t i = 0; i < n; ++i) {
    py::object it = items[i];
    const char* p;
    Py_ssize_t l;
    if (!object_bytes(it.ptr(), &p, &l))
      throw py::type_error("unhashable object column element");
    w(i) = (int64_t)murmur3_32(reinterpret_cast<const uint8_t*>(p), (size_t)l, seed);
  }
  return out;
}

struct SvHash {
  size_t operator()(const std::string_view& v) const noexcept { return std::hash<std::string_view>()(v); }
};

// Group equal keys (str compared by UTF-8 bytes, bytes raw; a column holds one
// kind): returns (first-occurrence row index per group, group id per row).
static py::tuple host_group_objects(py::sequence items) {
  const py::ssize_t n = py::len(items);
  std::unordered_map<std::string_view, int64_t, SvHash> ids;
  ids.reserve((size_t)n / 4 + 16);
  std::vector<py::object> keep;
  keep.reserve((size_t)n);
  std::vector<int64_t> first;
  py::array_t<int64_t> group(n);
  auto g = group.mutable_unchecked<1>();
  for (py::ssize_t i = 0; i < n; ++i) {
    py::object it = items[i];
    const char* p;
    Py_ssize_t l;
    if (!object_bytes(it.ptr(), &p, &l)) throw py::type_error("ungroupable object column element");
    keep.push_back(it);
    auto ins = ids.emplace(std::string_view(p, (size_t)l), (int64_t)first.size());
    if (ins.second) first.push_back((int64_t)i);
    g(i) = ins.first->second;
  }
  py::array_t<int64_t> firsts((py::ssize_t)first.size());
  if (!first.empty()) memcpy(firsts.mutable_data(), first.data(), first.size() * sizeof(int64_t));
  return py::make_tuple(firsts, group);
}

// ---------------------------------------------------------------------------
// Variable-length columns of the frame codec: [n+1 little-endian int64 offsets]
// [concatenated bytes].  str elements are UTF-8, bytes elements raw, None = empty.
static py::bytes codec_encode_varlen(py::sequence items, bool is_str) {
  const py::ssize_t n = py::len(items);
  std::vector<const char*> ptr((size_t)n);
  std::vector<py::ssize_t> len((size_t)n);
  std::vector<py::object> keep;  // temporaries (bytes(v) conversions) alive until copied
  size_t total = 0;
  for (py::ssize_t i = 0; i < n; ++i) {
    py::object it = items[i];
    PyObject* o = it.ptr();
    const char* p = "";
    py::ssize_t l = 0;
    if (o == Py_None) {
    } else if (is_str) {
      if (!PyUnicode_Check(o)) throw py::type_error("codec: string column holds a non-str value");
      p = PyUnicode_AsUTF8AndSize(o, &l);
      if (p == nullptr) throw py::error_already_set();
    } else {
      if (!PyBytes_Check(o)) {
        it = py::bytes(py::reinterpret_steal<py::object>(PyObject_Bytes(o)));
        if (!it) throw py::error_already_set();
        o = it.ptr();
      }
      char* q;
      if (PyBytes_AsStringAndSize(o, &q, &l) != 0) throw py::error_already_set();
      p = q;
    }
    keep.push_back(it);
    ptr[(size_t)i] = p;
    len[(size_t)i] = l;
    total += (size_t)l;
  }
  std::string out;
  out.resize((size_t)(n + 1) * 8 + total);
  int64_t off = 0;
  char* base = &out[0];
  char* data = base + (size_t)(n + 1) * 8;
  memcpy(base, &off, 8);
  for (py::ssize_t i = 0; i < n; ++i) {
    memcpy(data + off, ptr[(size_t)i], (size_t)len[(size_t)i]);
    off += len[(size_t)i];
    memcpy(base + (size_t)(i + 1) * 8, &off, 8);
  }
  return py::bytes(out);
}

// Returns a list of str / bytes; raises ValueError on inconsistent offsets.
static py::list codec_decode_varlen(py::buffer payload, int64_t n, bool is_str) {
  py::buffer_info info = payload.request();
  const size_t nbytes = (size_t)info.size * (size_t)info.itemsize;
  const char* base = static_cast<const char*>(info.ptr);
  const size_t ob = (size_t)(n + 1) * 8;
  if (n < 0 || nbytes < ob) throw py::value_error("short string column");
  const char* data = base + ob;
  const int64_t dlen = (int64_t)(nbytes - ob);
  py::list out((size_t)n);
  int64_t prev;
  memcpy(&prev, base, 8);
  if (n > 0 && prev != 0) throw py::value_error("corrupt string offsets");
  for (int64_t i = 0; i < n; ++i) {
    int64_t next;
    memcpy(&next, base + (size_t)(i + 1) * 8, 8);
    if (next < prev || next > dlen) throw py::value_error("corrupt string offsets");
    PyObject* o = is_str ? PyUnicode_DecodeUTF8(data + prev, (Py_ssize_t)(next - prev), nullptr)
                         : PyBytes_FromStringAndSize(data + prev, (Py_ssize_t)(next - prev));
    if (o == nullptr) throw py::error_already_set();
    PyList_SET_ITEM(out.ptr(), (Py_ssize_t)i, o);
    prev = next;
  }
  if (n > 0 && prev != dlen) throw py::value_error("corrupt string offsets");
  return out;
}

PYBIND11_MODULE(_bsnative, m) {
  m.doc() = "bigslice_b200 host runtime: evaluator core, crc32, murmur3";
  py::class_<EvalCore>(m, "EvalCore")
      .def(py::init<>())
      .def("set_count", &EvalCore::set_count)
      .def("has_group", &EvalCore::has_group)
      .def("add_group", &EvalCore::add_group)
      .def("wait_on", &EvalCore::wait_on)
      .def("schedule", &EvalCore::schedule)
      .def("done_ok", &EvalCore::done_ok)
      .def("take_todo", &EvalCore::take_todo)
      .def("num_waiting_groups", &EvalCore::num_waiting_groups);
  m.def("crc32", &py_crc32, py::arg("data"), py::arg("value") = 0);
  m.def("crc32_combine", &crc32_combine);
  m.def("crc32_combine_blocks", &py_crc32_combine_blocks);
  m.def("murmur3_strings", &py_murmur3_strings);
  m.def("murmur3_fixed", &py_murmur3_fixed);
  m.def("murmur3_bytes", &py_murmur3_bytes);
  m.def("murmur3_objects", &py_murmur3_objects);
  m.def("host_group_objects", &host_group_objects);
  m.def("codec_encode_varlen", &codec_encode_varlen);
  m.def("codec_decode_varlen", &codec_decode_varlen);
  m.def("host_hash_i64", &host_hash_i64, py::call_guard<py::gil_scoped_release>());
  m.def("host_partition_perm", &host_partition_perm, py::call_guard<py::gil_scoped_release>());
  m.def("host_hash_combine", &host_hash_combine, py::call_guard<py::gil_scoped_release>());
}
