// Host-side text parser for the reference's interaction files.
//
// Replaces the per-line Python loop of reference src/utils/IOUtil.py:7-16 (`loadSparseR`): lines are
// "u<sep>i" or "u<sep>i<sep>rating" with <sep> one of ',' ';' or whitespace (Util.py:5-11 `split_row`), CRLF tolerated;
// lines with another number of fields are skipped, exactly as the reference does.  One pass over an in-memory copy of the
// file; the caller (utils/IOUtil.py) builds the scipy matrix from the returned arrays.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {
inline bool is_sep(char c) { return c == ',' || c == ';' || c == ' ' || c == '\t' || c == '\r'; }
// the whole token must have been consumed (Python's int("3.0") / int("12abc") / float("4x") raise ValueError)
inline bool whole_token(const char* tok, const char* parsed_end, const char* eol) {
  return parsed_end != tok && (parsed_end >= eol || is_sep(*parsed_end));
}
}  // namespace

extern "C" int cf_parse_triplets(const char* path, int64_t* n_out, int64_t** users_out, int64_t** items_out, double** ratings_out) {
  CF_CHECK_ARG(path && n_out && users_out && items_out && ratings_out, "cf_parse_triplets: NULL argument");
  FILE* f = fopen(path, "rb");
  CF_CHECK_ARG(f != nullptr, "cf_parse_triplets: cannot open %s", path);
  fseek(f, 0, SEEK_END);
  const long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* buf = (char*)malloc((size_t)size + 2);
  if (!buf || fread(buf, 1, (size_t)size, f) != (size_t)size) {
    fclose(f);
    free(buf);
    cf_set_error("cf_parse_triplets: cannot read %s", path);
    return -1;
  }
  fclose(f);
  buf[size] = '\n';
  buf[size + 1] = 0;
  int64_t lines = 0;
  for (long k = 0; k <= size; ++k) lines += buf[k] == '\n';
  int64_t* us = (int64_t*)malloc(sizeof(int64_t) * (size_t)(lines + 1));
  int64_t* is = (int64_t*)malloc(sizeof(int64_t) * (size_t)(lines + 1));
  double* rs = (double*)malloc(sizeof(double) * (size_t)(lines + 1));
  if (!us || !is || !rs) {
    free(buf); free(us); free(is); free(rs);
    cf_set_error("cf_parse_triplets: out of host memory for %lld lines", (long long)lines);
    return -1;
  }
  int64_t n = 0;
  char* p = buf;
  char* end = buf + size + 1;
  while (p < end) {
    char* eol = (char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    // split the line into at most 4 fields
    const char* fld[4];
    int nf = 0;
    char* q = p;
    while (q < eol) {
      while (q < eol && is_sep(*q)) ++q;
      if (q >= eol) break;
      if (nf < 4) fld[nf] = q;
      ++nf;
      while (q < eol && !is_sep(*q)) ++q;
    }
    if (nf == 2 || nf == 3) {
      char* e1;
      char* e2;
      const long long u = strtoll(fld[0], &e1, 10);
      const long long i = strtoll(fld[1], &e2, 10);
      double r = 1.0;
      bool ok = whole_token(fld[0], e1, eol) && whole_token(fld[1], e2, eol);
      if (nf == 3) {
        char* e3;
        r = strtod(fld[2], &e3);
        ok = ok && whole_token(fld[2], e3, eol);
      }
      if (!ok) {
        free(buf); free(us); free(is); free(rs);
        cf_set_error("cf_parse_triplets: %s: malformed line %lld", path, (long long)n + 1);
        return -1;
      }
      us[n] = u;
      is[n] = i;
      rs[n] = r;
      ++n;
    }
    p = eol + 1;
  }
  free(buf);
  *n_out = n;
  *users_out = us;
  *items_out = is;
  *ratings_out = rs;
  return 0;
}

extern "C" void cf_free_host(void* p) { free(p); }
