// host_io.cpp -- host front-end of the gDCA path: the I/O that stays on the CPU.
//
//   gdca_read_fasta_alignment        DCAUtils read_fasta_alignment        (reference call site src/GaussDCA.jl:20)
//   gdca_remove_duplicate_sequences  DCAUtils remove_duplicate_sequences  (reference call site src/GaussDCA.jl:21-23)
//   gdca_write_rank / gdca_format_rank   printrank, "%i %i %e\n"          (reference src/GaussDCA.jl:67-74)
//
// Once the GPU path takes 0.08 s for a 200k-sequence alignment, parsing its 100 MB FASTA file dominates the wall
// clock of gDCA(filename): this reader does it in one pass over the mapped (or transparently gunzipped) bytes, encoding
// sequences in parallel (~0.07 s for 102 MB on 8 cores).  No CUDA here; these entry points work without a GPU.
//
// Semantics (identical to the oracle's reader, tests/test_host_cpu.py):
//   * a record starts at a line that begins with '>'; its sequence is the concatenation of the following lines
//     with white space removed;
//   * match columns = positions of the FIRST record whose character is not '.' and not a lowercase letter;
//     every other record must have the same length ("inputs are not aligned") and the same match columns
//     ("inconsistent inputs");
//   * a sequence is kept when (#'-' in match columns) / L <= max_gap_fraction;
//   * A C D E F G H I K L M N P Q R S T V W Y -> 1..20, everything else (B J O U X Z '-' ...) -> 21.
#include <cmath>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gdca_b200.h"

namespace {

thread_local std::string g_host_error;

int32_t host_fail(const char *msg) {
  g_host_error = msg;
  return GDCA_ERR_INVALID_ARG;
}

struct Tables {
  int8_t code[256];
  bool match[256];
  Tables() {
    for (int c = 0; c < 256; ++c) {
      code[c] = 21;
      match[c] = !(c == '.' || (c >= 'a' && c <= 'z'));
    }
    const char *aa = "ACDEFGHIKLMNPQRSTVWY";
    for (int k = 0; aa[k]; ++k) code[(unsigned char)aa[k]] = (int8_t)(k + 1);
  }
};
const Tables T;

// The bytes of a file: plain files are mapped (no copy, no zero-filled staging buffer: the single pass of the encoder is the
// only time the pages are touched); gzip files (magic 1f 8b) are inflated with zlib into an owned buffer.
struct FileBytes {
  const char *p = nullptr;
  size_t n = 0;
  std::vector<char> own;
  void *map = nullptr;
  size_t map_len = 0;
  ~FileBytes() {
    if (map) munmap(map, map_len);
  }
};

bool read_all(const char *path, FileBytes &fb) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return false;
  unsigned char magic[2] = {0, 0};
  const ssize_t got2 = pread(fd, magic, 2, 0);
  if (!(got2 == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
    struct stat st;
    if (fstat(fd, &st) != 0) {
      close(fd);
      return false;
    }
    if (st.st_size > 0 && S_ISREG(st.st_mode)) {
      void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);  // one bulk prefault
      if (m != MAP_FAILED) {
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        fb.map = m;
        fb.map_len = (size_t)st.st_size;
        fb.p = (const char *)m;
        fb.n = (size_t)st.st_size;
        close(fd);
        return true;
      }
    }
    // not mappable (pipe, empty file, ...): plain reads
    fb.own.clear();
    char chunk[1 << 16];
    for (;;) {
      const ssize_t rd = read(fd, chunk, sizeof chunk);
      if (rd < 0) {
        close(fd);
        return false;
      }
      if (rd == 0) break;
      fb.own.insert(fb.own.end(), chunk, chunk + rd);
    }
    close(fd);
    fb.p = fb.own.data();
    fb.n = fb.own.size();
    return true;
  }
  close(fd);
  std::vector<char> &buf = fb.own;
  gzFile f = gzopen(path, "rb");
  if (!f) return false;
  gzbuffer(f, 1 << 20);
  size_t used = 0;
  buf.resize(1 << 22);
  for (;;) {
    if (used == buf.size()) buf.resize(buf.size() * 2);
    const size_t want = buf.size() - used;
    const int got = gzread(f, buf.data() + used, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
    if (got < 0) {
      gzclose(f);
      return false;
    }
    if (got == 0) break;
    used += (size_t)got;
  }
  gzclose(f);
  buf.resize(used);
  fb.p = buf.data();
  fb.n = used;
  return true;
}

struct Rec {
  size_t beg, end;  // sequence bytes (may contain line breaks / blanks) in the file buffer
};

inline bool is_ws(char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }

// 64-bit FNV-1a over a row; good enough for a verified (memcmp) hash set
inline uint64_t row_hash(const int8_t *p, int64_t n) {
  uint64_t h = 1469598103934665603ull;
  for (int64_t i = 0; i < n; ++i) {
    h ^= (uint8_t)p[i];
    h *= 1099511628211ull;
  }
  return h ^ (h >> 29);
}

}  // namespace

extern "C" {

const char *gdca_host_last_error(void) { return g_host_error.c_str(); }

void gdca_free_host(void *p) { free(p); }

int32_t gdca_read_fasta_alignment(const char *path, double max_gap_fraction, int8_t **Z_out, int64_t *L_out,
                                  int64_t *M_out) {
  if (!path || !Z_out || !L_out || !M_out) return host_fail("read_fasta_alignment: NULL argument");
  *Z_out = nullptr;
  *L_out = *M_out = 0;
  FileBytes fb;
  if (!read_all(path, fb)) {
    g_host_error = std::string("cannot open file ") + path;
    return GDCA_ERR_INVALID_ARG;
  }
  // ---- index the records (a header is a line starting with '>')
  std::vector<Rec> recs;
  const char *buf = fb.p;
  const size_t n = fb.n;
  size_t pos = 0;
  bool in_record = false;
  while (pos < n) {
    const char *nl = (const char *)memchr(buf + pos, '\n', n - pos);
    const size_t eol = nl ? (size_t)(nl - buf) : n;
    if (buf[pos] == '>') {
      if (in_record) recs.back().end = pos;
      recs.push_back(Rec{eol < n ? eol + 1 : n, n});
      in_record = true;
    }
    pos = eol + 1;
  }
  if (recs.empty()) return host_fail("no sequences found in the FASTA file");

  // ---- first record: length and match columns
  auto seq_len = [&](const Rec &r) {
    size_t len = 0;
    for (size_t p = r.beg; p < r.end; ++p) len += !is_ws(buf[p]);
    return len;
  };
  const size_t len0 = seq_len(recs[0]);
  std::vector<uint8_t> is_match(len0);
  int64_t L = 0;
  {
    size_t c = 0;
    for (size_t p = recs[0].beg; p < recs[0].end; ++p) {
      if (is_ws(buf[p])) continue;
      is_match[c] = T.match[(unsigned char)buf[p]];
      L += is_match[c];
      ++c;
    }
  }
  if (L == 0) return host_fail("alignment has no match columns");

  // ---- every record: validate, count gaps, encode (parallel over records)
  const int64_t R = (int64_t)recs.size();
  // 2 MB-aligned and advised for huge pages: the parallel first touch of a 100 MB result is otherwise 25k page faults
  int8_t *all = nullptr;
  {
    void *mem = nullptr;
    const size_t bytes = (size_t)R * (size_t)L;
    if (posix_memalign(&mem, (size_t)2 << 20, bytes ? bytes : 1) != 0) return host_fail("out of host memory");
    madvise(mem, bytes, MADV_HUGEPAGE);
    all = (int8_t *)mem;
  }
  std::vector<uint8_t> keep((size_t)R, 0);
  int bad = 0;  // 1: not aligned, 2: inconsistent
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < R; ++r) {
    int8_t *row = all + (size_t)r * (size_t)L;
    size_t c = 0;
    int64_t m = 0, gaps = 0;
    int err = 0;
    for (size_t p = recs[r].beg; p < recs[r].end; ++p) {
      const char ch = buf[p];
      if (is_ws(ch)) continue;
      if (c >= len0) {
        err = 1;
        break;
      }
      const bool mt = T.match[(unsigned char)ch];
      if (mt != (bool)is_match[c]) {
        err = 2;
        // keep scanning to tell "not aligned" (length) from "inconsistent" (columns) like the reference does
      } else if (mt) {
        row[m++] = T.code[(unsigned char)ch];
        gaps += (ch == '-');
      }
      ++c;
    }
    if (!err && c != len0) err = 1;
    if (err == 2 && c != len0) err = 1;
    if (err) {
#pragma omp critical
      if (!bad || err < bad) bad = err;
      continue;
    }
    keep[(size_t)r] = ((double)gaps / (double)L <= max_gap_fraction);
  }
  if (bad) {
    free(all);
    return host_fail(bad == 1 ? "inputs are not aligned" : "inconsistent inputs");
  }
  // ---- compact the kept sequences, order preserved
  int64_t M = 0;
  for (int64_t r = 0; r < R; ++r) {
    if (!keep[(size_t)r]) continue;
    if (M != r) memmove(all + (size_t)M * (size_t)L, all + (size_t)r * (size_t)L, (size_t)L);
    ++M;
  }
  if (M == 0) {
    free(all);
    char b[160];
    snprintf(b, sizeof b, "Out of %lld sequences, none passed the filter (max_gap_fraction=%g)", (long long)R, max_gap_fraction);
    return host_fail(b);
  }
  *Z_out = all;
  *L_out = L;
  *M_out = M;
  return GDCA_OK;
}

int32_t gdca_remove_duplicate_sequences(const int8_t *Z, int64_t L, int64_t M, int8_t *Z_out, int64_t *M_out,
                                        int64_t *kept) {
  if (!Z || !Z_out || !M_out || L < 1 || M < 1) return host_fail("remove_duplicate_sequences: bad argument");
  size_t cap = 16;
  while (cap < (size_t)M * 2) cap <<= 1;
  std::vector<int64_t> table(cap, -1);
  // the row hashes are the bulk of the work (every byte once): computed in parallel; the insertion below stays serial
  // because "keep the FIRST occurrence, in order" is inherently sequential
  std::vector<uint64_t> hashes((size_t)M);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < M; ++k) hashes[(size_t)k] = row_hash(Z + (size_t)k * (size_t)L, L);
  int64_t out = 0;
  for (int64_t k = 0; k < M; ++k) {
    const int8_t *row = Z + (size_t)k * (size_t)L;
    size_t h = (size_t)hashes[(size_t)k] & (cap - 1);
    bool dup = false;
    while (table[h] >= 0) {
      if (memcmp(Z_out + (size_t)table[h] * (size_t)L, row, (size_t)L) == 0) {
        dup = true;
        break;
      }
      h = (h + 1) & (cap - 1);
    }
    if (dup) continue;
    table[h] = out;
    if (Z_out + (size_t)out * (size_t)L != row) memmove(Z_out + (size_t)out * (size_t)L, row, (size_t)L);
    if (kept) kept[out] = k;
    ++out;
  }
  *M_out = out;
  return GDCA_OK;
}

// printrank(outfile, R): one "%i %i %e\n" line per row (src/GaussDCA.jl:67-74)
// Rows are formatted in parallel (one chunk of rows per OpenMP thread, "%lld %lld %e\n" -- at most 64 bytes per row) and the
// chunks are written in order: the bytes are exactly those of the serial loop of printrank (src/GaussDCA.jl:67-74).  At
// L = 1500 the ranking has 1.1 M rows: 0.38 s with a serial fprintf loop (5x the whole GPU path at L = 500), 0.08 - 0.17 s here
// on 8 warm cores.
// one row, Julia's @printf spelling of the non-finite values: "NaN", "Inf", "-Inf" (C's %e prints nan / inf / -inf)
static int format_row(char *out, const gdca_rank_t &r) {
  if (std::isfinite(r.score)) return snprintf(out, 64, "%lld %lld %e\n", (long long)r.i, (long long)r.j, r.score);
  const char *v = std::isnan(r.score) ? "NaN" : (r.score > 0 ? "Inf" : "-Inf");
  return snprintf(out, 64, "%lld %lld %s\n", (long long)r.i, (long long)r.j, v);
}

static int64_t format_rows(const gdca_rank_t *R, int64_t lo, int64_t hi, char *out) {
  int64_t off = 0;
  for (int64_t k = lo; k < hi; ++k)
    off += format_row(out + off, R[k]);
  return off;
}

int32_t gdca_write_rank(const char *path, const gdca_rank_t *R, int64_t n) {
  if (!path || (!R && n > 0) || n < 0) return host_fail("write_rank: bad argument");
  FILE *f = fopen(path, "w");
  if (!f) {
    g_host_error = std::string("cannot open file ") + path;
    return GDCA_ERR_INVALID_ARG;
  }
  constexpr int64_t CH = 1 << 12;   // rows per chunk: 256 KB of buffer, reused wave after wave (stays in cache)
  constexpr int64_t WAVE = 32;      // chunks formatted per parallel wave
  const int64_t nchunks = (n + CH - 1) / CH;
  std::vector<std::vector<char>> bufs((size_t)(nchunks < WAVE ? nchunks : WAVE));
  std::vector<int64_t> lens(bufs.size());
  bool ok = true;
  for (int64_t c0 = 0; c0 < nchunks && ok; c0 += WAVE) {
    const int64_t c1 = c0 + WAVE < nchunks ? c0 + WAVE : nchunks;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = c0; c < c1; ++c) {
      const int64_t lo = c * CH, hi = lo + CH < n ? lo + CH : n;
      std::vector<char> &b = bufs[(size_t)(c - c0)];
      if (b.size() < (size_t)CH * 64 + 1) b.resize((size_t)CH * 64 + 1);
      lens[(size_t)(c - c0)] = format_rows(R, lo, hi, b.data());
    }
    for (int64_t c = c0; c < c1 && ok; ++c)
      ok = fwrite(bufs[(size_t)(c - c0)].data(), 1, (size_t)lens[(size_t)(c - c0)], f) == (size_t)lens[(size_t)(c - c0)];
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? GDCA_OK : host_fail("write_rank: write failed");
}

// printrank(io, R) for hosts that own the stream: formats into buf; returns the bytes needed in *used
// (call with cap = 0 to size the buffer: at most 64 bytes per row).  Two parallel passes: sizes, then bytes in place.
int32_t gdca_format_rank(const gdca_rank_t *R, int64_t n, char *buf, int64_t cap, int64_t *used) {
  if ((!R && n > 0) || n < 0 || !used) return host_fail("format_rank: bad argument");
  constexpr int64_t CH = 1 << 12;
  const int64_t nchunks = (n + CH - 1) / CH;
  std::vector<int64_t> off((size_t)nchunks + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t lo = c * CH, hi = lo + CH < n ? lo + CH : n;
    char tmp[96];
    int64_t len = 0;
    for (int64_t k = lo; k < hi; ++k)
      len += format_row(tmp, R[k]);
    off[(size_t)c + 1] = len;
  }
  for (int64_t c = 0; c < nchunks; ++c) off[(size_t)c + 1] += off[(size_t)c];
  const int64_t total = off[(size_t)nchunks];
  *used = total;
  if (!buf) return GDCA_OK;
  if (total > cap) return host_fail("format_rank: buffer too small");
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t lo = c * CH, hi = lo + CH < n ? lo + CH : n;
    std::vector<char> b((size_t)(hi - lo) * 64 + 1);   // snprintf writes a trailing NUL: format aside, then copy
    const int64_t len = format_rows(R, lo, hi, b.data());
    memcpy(buf + off[(size_t)c], b.data(), (size_t)len);
  }
  return GDCA_OK;
}

}  // extern "C"
