// Host-side reader of the reference's dataset text format (utility_data/data_loader.py:48-70): one line per user,
// "user item item ...".  Replaces the per-line Python loop (4.4 s on the yelp2018 shape, minutes at the scale-up
// shape) by one pass over the file image.  Host code: the arrays it fills are what Data hands to the device CSR
// builders (idg_csr_structure) -- same order (file order), same ids, same max-id rule as the reference.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "idg_common.cuh"

using namespace idg;

// Two-call protocol: call with pair_cap = line_cap = 0 to count (*n_pairs, *n_lines), allocate, call again.
//   h_user/h_item [n_pairs]   one entry per (user, item) token pair, file order          (inter_users / inter_items)
//   h_line_user   [n_lines]   first token of every non-blank line, also lines without items (unique_users)
//   h_line_len    [n_lines]   number of items on that line (0 for a user with an empty line; pos_length keeps the > 0 ones)
//   *max_user / *max_item     maxima over lines that have at least one item (data_loader.py:62-63); -1 if none
// Tokens are separated by blanks/tabs; a character that is neither a digit nor white space is an error (-4), as
// int() raises in the reference.
extern "C" int idg_parse_ratings(const char* path, int64_t* h_user, int64_t* h_item, int64_t pair_cap, int64_t* n_pairs,
                                 int64_t* h_line_user, int64_t* h_line_len, int64_t line_cap, int64_t* n_lines, int64_t* max_user,
                                 int64_t* max_item) {
    if (!path || !n_pairs || !n_lines || !max_user || !max_item) return fail(-1, "idg_parse_ratings: bad argument%s");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(-3, "idg_parse_ratings: cannot open %s", path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)(sz > 0 ? sz : 0) + 1);
    const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    buf[got] = '\n';
    const char* p = buf.data();
    const char* end = p + got;
    int64_t np = 0, nl = 0, mu = -1, mi = -1;
    const bool fill = (pair_cap > 0 || line_cap > 0);
    while (p < end) {
        // one line
        int64_t user = 0, n_items = 0, line_max = -1;
        bool have_user = false;
        const int64_t first_pair = np;
        while (p < end && *p != '\n') {
            const char c = *p;
            if (c == ' ' || c == '\t' || c == '\r') { ++p; continue; }
            bool neg = false;
            if (c == '-') { neg = true; ++p; }
            if (p >= end || *p < '0' || *p > '9') return fail(-4, "idg_parse_ratings: %s: not an integer token at byte %lld", path, (long long)(p - buf.data()));
            int64_t v = 0;
            while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
            if (neg) v = -v;
            if (p < end && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n'))
                return fail(-4, "idg_parse_ratings: %s: not an integer token at byte %lld", path, (long long)(p - buf.data()));
            if (!have_user) { user = v; have_user = true; continue; }
            if (fill) {
                if (np >= pair_cap) return fail(-2, "idg_parse_ratings: pair buffer too small%s");
                h_user[np] = user; h_item[np] = v;
            }
            ++np; ++n_items;
            if (v > line_max) line_max = v;
        }
        if (p < end) ++p;  // the newline
        if (!have_user) continue;  // blank line
        (void)first_pair;
        if (fill) {
            if (nl >= line_cap) return fail(-2, "idg_parse_ratings: line buffer too small%s");
            h_line_user[nl] = user; h_line_len[nl] = n_items;
        }
        ++nl;
        if (n_items > 0) {
            if (user > mu) mu = user;
            if (line_max > mi) mi = line_max;
        }
    }
    *n_pairs = np; *n_lines = nl; *max_user = mu; *max_item = mi;
    return 0;
}
