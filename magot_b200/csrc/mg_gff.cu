// mg_gff.cu -- native host reader + flattener for GFF3 / GTF annotation text (host code only).
//
// Replaces the per-line Python object construction of read_gff (genome.py:242-415) and the per-object walk of
// ParentAnnotation.get_fasta / AnnotationSet.get_fasta (genome.py:578-582, :677-731) on the path
//     annotation text -> SoA interval tables (contig id, start, end, strand, phase per segment, sorted per transcript)
// that feeds K1/K2/K3.  The reference spends ~1 ms per line there (AnnotationSet.__getitem__ evaluates every dict attribute,
// genome.py:536-544); this reader interns every string once and does the ID / de-dup / implicit-parent / child-list semantics
// on integer ids:
//   * a line is accepted iff it does not start with '#' and holds exactly 8 tabs (genome.py:283); '\r' and the
//     features_to_replace pairs are applied to the accepted line (:285-286) before the fields are split;
//   * version: '=' in column 9 -> GFF3, else GFF2 with the gene_id / transcript_id probing of :288-300;
//   * attributes: GFF3 key=value (text between the first two '='), GFF2 key "value" / key value (:321-333);
//   * Parent from parent_field, else the first key of parents_hierarchy that is present (:335-342);
//   * ID from IDfield, else parent-featuretype, else seqid-featuretypeSTART (:344-353);
//   * de-dup (:355-364): an ID that any table already holds becomes ID2 the first time, ID-3, ID-4, ... afterwards (the renamed
//     ID is not re-checked); "holds" is AnnotationSet.__getitem__: the table whose NAME sorts last wins;
//   * implicit parents from the hierarchy keys (:366-388), child lists without duplicates (:390-392), a missing parent stops
//     the read (:393-399);
//   * base_features -> BaseAnnotation rows, everything else -> ParentAnnotation rows, filed under the feature type (:405-413).
// The Python side (magot_b200/gffnative.py) turns the model into the reference's objects only when somebody asks for them;
// gff2fasta goes from here straight to a RecordTable (mg_gff_flatten: which children, in which order, under which header --
// genome.py:683-719 -- on the integer model).
#include <stdint.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <deque>
#include <string>
#include <vector>
#include "../../include/magot_b200.h"

void mg_set_error(const char *fmt, ...);

namespace {

struct Str { const char *p; uint32_t len; uint32_t hash; };

static inline uint32_t fnv1a(const char *p, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)p[i]; h *= 16777619u; }
    return h;
}

// string interner: open addressing, ids are dense int32
struct Interner {
    std::vector<Str> strs;
    std::vector<int32_t> slots;
    std::deque<std::string> arena;                   // constructed strings (stable addresses)
    size_t mask = 0;
    Interner() { slots.assign(1 << 16, -1); mask = slots.size() - 1; }
    void grow() {
        std::vector<int32_t> ns(slots.size() * 4, -1);
        const size_t m = ns.size() - 1;
        for (size_t i = 0; i < strs.size(); i++) {
            size_t s = strs[i].hash & m;
            while (ns[s] >= 0) s = (s + 1) & m;
            ns[s] = (int32_t)i;
        }
        slots.swap(ns);
        mask = m;
    }
    int32_t find(const char *p, size_t n, uint32_t h) const {
        size_t s = h & mask;
        while (true) {
            const int32_t id = slots[s];
            if (id < 0) return -1;
            const Str &t = strs[id];
            if (t.hash == h && t.len == n && memcmp(t.p, p, n) == 0) return id;
            s = (s + 1) & mask;
        }
    }
    // `stable`: p stays valid for the life of the model (it points into the caller's text)
    int32_t intern(const char *p, size_t n, bool stable) {
        const uint32_t h = fnv1a(p, n);
        const int32_t f = find(p, n, h);
        if (f >= 0) return f;
        if (!stable) { arena.emplace_back(p, n); p = arena.back().data(); }
        if ((strs.size() + 1) * 2 > slots.size()) grow();
        size_t s = h & mask;
        while (slots[s] >= 0) s = (s + 1) & mask;
        slots[s] = (int32_t)strs.size();
        strs.push_back({p, (uint32_t)n, h});
        return (int32_t)strs.size() - 1;
    }
    int32_t intern(const std::string &s) { return intern(s.data(), s.size(), false); }
};

// 64-bit key -> int64 value map (open addressing; key 0 is not used by the callers: keys are shifted by one)
struct Map64 {
    std::vector<uint64_t> keys;
    std::vector<int64_t> vals;
    size_t used = 0, mask = 0;
    Map64() { keys.assign(1 << 12, 0); vals.assign(1 << 12, 0); mask = keys.size() - 1; }
    static inline uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
    void grow() {
        std::vector<uint64_t> nk(keys.size() * 4, 0);
        std::vector<int64_t> nv(keys.size() * 4, 0);
        const size_t m = nk.size() - 1;
        for (size_t i = 0; i < keys.size(); i++) if (keys[i]) {
            size_t s = mix(keys[i]) & m;
            while (nk[s]) s = (s + 1) & m;
            nk[s] = keys[i]; nv[s] = vals[i];
        }
        keys.swap(nk); vals.swap(nv); mask = m;
    }
    int64_t *get(uint64_t k) {
        k += 1;
        size_t s = mix(k) & mask;
        while (keys[s]) { if (keys[s] == k) return &vals[s]; s = (s + 1) & mask; }
        return nullptr;
    }
    int64_t *put(uint64_t k, int64_t v) {               // inserts or overwrites; returns the slot
        if ((used + 1) * 2 > keys.size()) grow();
        k += 1;
        size_t s = mix(k) & mask;
        while (keys[s]) { if (keys[s] == k) { vals[s] = v; return &vals[s]; } s = (s + 1) & mask; }
        keys[s] = k; vals[s] = v; used++;
        return &vals[s];
    }
};

struct Row {
    int32_t id, seqid, ftype, strand, source, parent;   // string ids (-1 = None)
    int64_t start, end;
    double score;
    int8_t has_score, phase, is_base, implicit;
    int32_t ext;                                        // >= 0: object that already existed in the set (index given by the caller)
    int32_t table;                                      // table (dict attribute) index
    int64_t attr0; int32_t nattr;                       // defline attributes (IDfield / parent_field keys removed), in defline order
    int64_t start_text;                                 // unused placeholder for alignment
    int64_t child_head, child_tail;                     // child IDs in append order: linked list in mg_gff::kid_id / kid_next
    int32_t n_children;
    int32_t ext_children0;                              // children an ext row already had (only the later ones are new)
};

struct Table { int32_t name; std::vector<int64_t> rows; bool preexisting; };   // rows in dict insertion order

enum { ST_OK = 0, ST_GFF2_NO_VALUE = 1, ST_MISSING_PARENT = 2, ST_GFF3_NO_EQ = 3, ST_BAD_INT = 4, ST_GFF2_NO_KEY = 5,
       ST_PARENT_IS_BASE = 6, ST_ATTR_CLASH = 7, ST_FEW_FIELDS = 8, ST_NONE_TWICE = 9 };

static inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\x0b' || c == '\x0c'; }

struct Opts {
    int version = 0;
    std::vector<std::string> ignore, base, hier;
    bool ignore_is_str = false, base_is_str = false;
    std::vector<std::pair<std::string, std::string>> replace;
    bool has_id_field = true, has_parent_field = true;
    std::string id_field = "ID", parent_field = "Parent";
};

}  // namespace

struct mg_gff {
    Interner in;
    std::vector<Row> rows;
    std::vector<int32_t> attr_key, attr_val;
    std::vector<Table> tables;
    Map64 table_of_name;                                // ftype string id -> table index
    std::vector<int32_t> nondict;                       // attribute names of the set that are not tables
    Map64 child_seen;                                   // (row << 32 | child id) -> 1
    std::vector<int32_t> kid_id;                        // child lists: one node per appended child
    std::vector<int64_t> kid_next;
    std::vector<int64_t> owner_row;                     // owner, indexed by string id (-1: no table holds the ID)
    // which table files an ID, and where: (table << 40 | position) of the FIRST table that filed it, indexed by string id
    // (sequential ids: sequential memory); the few IDs that a second table files as well live in `filed_more`
    std::vector<int64_t> filed_first;
    Map64 filed_more;                                   // (table << 32 | id) -> position
    // the one parent row a child ID has been appended to (-1 none yet, -2 several: see child_seen), indexed by string id
    std::vector<int64_t> child_parent;
    std::vector<int32_t> newid_cnt;                     // generate_new_ID_dict, indexed by string id (0: not in the dict)
    Opts o;
    int32_t none_id = -1;                               // string id that stands for Python's None as an ID
    int64_t n_lines = 0;                                // accepted, non-ignored lines
    int status = ST_OK;
    int64_t err_a = -1, err_b = -1, err_c = -1;         // string ids for the caller's message
    std::vector<char> keep;                             // cleaned copies of lines that held '\r' or replacements (string storage)
    // exported, flat copies (built by finish())
    std::vector<int32_t> c_id, c_seqid, c_ftype, c_strand, c_source, c_parent, c_ext, c_table, c_nattr, c_ext_children0;
    std::vector<int64_t> c_start, c_end, c_attr0, c_child_off;
    std::vector<double> c_score;
    std::vector<int8_t> c_has_score, c_phase, c_is_base, c_implicit;
    std::vector<int32_t> c_child;
    std::vector<int64_t> c_str_off;
    std::vector<char> c_pool;
    std::vector<int32_t> c_table_name;
    std::vector<int64_t> c_table_off, c_table_rows;

    int table_index(int32_t name) {
        int64_t *t = table_of_name.get((uint64_t)name);
        return t ? (int)*t : -1;
    }
    int make_table(int32_t name, bool pre) {
        tables.emplace_back();
        tables.back().name = name;
        tables.back().preexisting = pre;
        table_of_name.put((uint64_t)name, (int64_t)tables.size() - 1);
        return (int)tables.size() - 1;
    }
    // adict[feature_type] on demand; a non-dict attribute of that name makes the reference fail
    int table_for(int32_t name) {
        int t = table_index(name);
        if (t >= 0) return t;
        for (int32_t nd : nondict) if (nd == name) return -2;
        return make_table(name, false);
    }
    // register(): owner[ID] = row unless a table whose name sorts later already holds the ID
    void reg(int32_t id, int64_t row) {
        int64_t *cur = owner_get(id);
        if (cur) {
            const Str &a = in.strs[tables[rows[row].table].name], &b = in.strs[tables[rows[*cur].table].name];
            const int c = memcmp(a.p, b.p, a.len < b.len ? a.len : b.len);
            const bool ge = c > 0 || (c == 0 && a.len >= b.len);
            if (!ge) return;
        }
        owner_put(id, row);
    }
    int64_t *filed_get(int t, int32_t id) {
        if ((size_t)id < filed_first.size() && filed_first[id] >= 0) {
            if ((filed_first[id] >> 40) == t) { filed_tmp = filed_first[id] & ((1ll << 40) - 1); return &filed_tmp; }
            return filed_more.get(((uint64_t)t << 32) | (uint32_t)id);
        }
        return nullptr;
    }
    void filed_put(int t, int32_t id, int64_t posn) {
        if ((size_t)id >= filed_first.size()) filed_first.resize(std::max((size_t)id + 1, filed_first.size() * 2 + 1024), -1);
        if (filed_first[id] < 0) filed_first[id] = ((int64_t)t << 40) | posn;
        else filed_more.put(((uint64_t)t << 32) | (uint32_t)id, posn);
    }
    int64_t filed_tmp = 0;
    bool add_child(int64_t row, int32_t child) {
        Row &r = rows[row];
        if (r.child_tail >= 0 && kid_id[r.child_tail] == child) return false;      // the usual repeat: the same child as last time
        if ((size_t)child >= child_parent.size()) child_parent.resize(std::max((size_t)child + 1, child_parent.size() * 2 + 1024), -1);
        const uint64_t k = ((uint64_t)row << 32) | (uint32_t)child;
        int64_t &cp = child_parent[child];
        if (cp == -1) cp = row;                            // first parent of this child: no set needed
        else if (cp >= 0) {
            if (cp == row) return false;
            child_seen.put(((uint64_t)cp << 32) | (uint32_t)child, 1);      // a second parent: from now on the exact set
            child_seen.put(k, 1);
            cp = -2;
        } else {
            if (child_seen.get(k)) return false;
            child_seen.put(k, 1);
        }
        kid_id.push_back(child);
        kid_next.push_back(-1);
        const int64_t node = (int64_t)kid_id.size() - 1;
        if (r.child_tail >= 0) kid_next[r.child_tail] = node; else r.child_head = node;
        r.child_tail = node;
        r.n_children++;
        return true;
    }
    int64_t *owner_get(int32_t id) {
        if ((size_t)id >= owner_row.size() || owner_row[id] < 0) return nullptr;
        return &owner_row[id];
    }
    void owner_put(int32_t id, int64_t row) {
        if ((size_t)id >= owner_row.size()) owner_row.resize(std::max((size_t)id + 1, owner_row.size() * 2 + 1024), -1);
        owner_row[id] = row;
    }
    int64_t new_row() {
        rows.emplace_back();
        Row &r = rows.back();
        r.id = r.seqid = r.ftype = r.strand = r.source = r.parent = -1;
        r.start = r.end = 0; r.score = 0; r.has_score = 0; r.phase = -1; r.is_base = 0; r.implicit = 0; r.ext = -1; r.table = -1;
        r.attr0 = (int64_t)attr_key.size(); r.nattr = 0; r.ext_children0 = 0; r.start_text = 0;
        r.child_head = r.child_tail = -1; r.n_children = 0;
        return (int64_t)rows.size() - 1;
    }
    // adict[table][ID] = row (a repeated key keeps its place in the dict)
    void file_row(int t, int32_t id, int64_t row) {
        int64_t *slot = filed_get(t, id);
        if (slot) { tables[t].rows[*slot] = row; }
        else { filed_put(t, id, (int64_t)tables[t].rows.size()); tables[t].rows.push_back(row); }
    }
};

namespace {

// sequential reader of the option blob written by gffnative.py: int32 / length-prefixed strings
struct Blob {
    const uint8_t *p, *e;
    bool ok = true;
    int32_t i32() { if (p + 4 > e) { ok = false; return 0; } int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    std::string str() { const int32_t n = i32(); if (n < 0 || p + n > e) { ok = false; return std::string(); } std::string s((const char *)p, n); p += n; return s; }
};

static bool in_list(const std::vector<std::string> &v, bool is_str, const char *p, size_t n) {
    if (is_str) {                                       // `feature_type in "some string"`: substring test
        if (v.empty()) return false;
        const std::string &s = v[0];
        if (n == 0) return true;
        return s.size() >= n && memmem(s.data(), s.size(), p, n) != nullptr;
    }
    for (const std::string &s : v) if (s.size() == n && memcmp(s.data(), p, n) == 0) return true;
    return false;
}

// int(text) as Python 2 accepts it here: optional surrounding white space, optional sign, decimal digits
static bool parse_int(const char *p, size_t n, int64_t *out) {
    size_t a = 0, b = n;
    while (a < b && is_ws(p[a])) a++;
    while (b > a && is_ws(p[b - 1])) b--;
    if (a == b) return false;
    bool neg = false;
    if (p[a] == '+' || p[a] == '-') { neg = p[a] == '-'; a++; }
    if (a == b || b - a > 18) return false;
    int64_t v = 0;
    for (size_t i = a; i < b; i++) { if (p[i] < '0' || p[i] > '9') return false; v = v * 10 + (p[i] - '0'); }
    *out = neg ? -v : v;
    return true;
}

// float(text): strtod on the stripped token, whole token consumed; hex floats / underscores are not Python-2 floats
static bool parse_float(const char *p, size_t n, double *out) {
    size_t a = 0, b = n;
    while (a < b && is_ws(p[a])) a++;
    while (b > a && is_ws(p[b - 1])) b--;
    if (a == b || b - a > 62) return false;
    char buf[64];
    memcpy(buf, p + a, b - a);
    buf[b - a] = 0;
    for (size_t i = 0; i < b - a; i++) if (buf[i] == 'x' || buf[i] == 'X' || buf[i] == '_' || buf[i] == '(' || buf[i] == 'p' || buf[i] == 'P') return false;
    char *end = nullptr;
    const double v = strtod(buf, &end);
    if (end != buf + (b - a)) return false;
    *out = v;
    return true;
}

struct KV { const char *k; size_t kn; const char *v; size_t vn; };

}  // namespace

// ---- the reader ---------------------------------------------------------------------------------------------------------
// opts: blob of gffnative.py (_pack_opts).  `text` must stay alive and unchanged while the handle lives (strings point into it).
extern "C" int mg_gff_parse(const uint8_t *text, int64_t n, const uint8_t *opts, int64_t n_opts, mg_gff **out) {
    if (!out || (!text && n > 0) || !opts) { mg_set_error("mg_gff_parse: NULL argument"); return MG_EINVAL; }
    mg_gff *m = new mg_gff();
    Blob b{opts, opts + n_opts};
    Opts &o = m->o;
    o.version = b.i32();
    o.ignore_is_str = b.i32() != 0;
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) o.ignore.push_back(b.str());
    o.base_is_str = b.i32() != 0;
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) o.base.push_back(b.str());
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) o.hier.push_back(b.str());
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) { std::string f = b.str(), t = b.str(); o.replace.emplace_back(f, t); }
    o.has_id_field = b.i32() != 0;
    o.id_field = b.str();
    o.has_parent_field = b.i32() != 0;
    o.parent_field = b.str();
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) m->make_table(m->in.intern(b.str()), true);      // existing tables, __dict__ order
    for (int32_t k = b.i32(); k > 0 && b.ok; k--) m->nondict.push_back(m->in.intern(b.str()));
    for (int32_t k = b.i32(), x = 0; k > 0 && b.ok; k--, x++) {                                    // existing objects
        const int32_t t = b.i32();
        const int32_t id = m->in.intern(b.str());
        const int32_t is_parent = b.i32();
        const int64_t row = m->new_row();
        Row &r = m->rows[row];
        r.id = id; r.ext = x; r.table = t; r.is_base = is_parent ? 0 : 1; r.ftype = m->tables[t].name;
        for (int32_t c = b.i32(); c > 0 && b.ok; c--) m->add_child(row, m->in.intern(b.str()));
        m->rows[row].ext_children0 = m->rows[row].n_children;
        m->file_row(t, id, row);
    }
    if (!b.ok) { delete m; mg_set_error("mg_gff_parse: malformed option blob"); return MG_EINVAL; }
    // owner map of the existing objects: tables in sorted-name order, later names win (AnnotationSet.__getitem__)
    for (size_t r = 0; r < m->rows.size(); r++) m->reg(m->rows[r].id, (int64_t)r);

    const char *T = (const char *)text;
    const int32_t none_id = m->in.intern("\0\0None\0\0", 8, false);      // stands for the key None (cannot occur in a line: no NUL... by convention)
    m->none_id = none_id;
    std::vector<KV> kv;
    std::vector<std::pair<const char *, size_t>> f;
    std::string clean, tmp, scratch;
    // consecutive lines repeat most of their strings (seqid, source, type, strand, attribute keys, transcript / gene ids):
    // a string equal to the one last seen in the same role keeps its id without being hashed
    struct Last { const char *p = nullptr; size_t n = 0; int32_t id = -1; };
    Last last_role[12];
    auto intern_role = [&](int role, const char *p, size_t n2, bool st) -> int32_t {
        Last &l = last_role[role];
        if (l.id >= 0 && l.n == n2 && memcmp(l.p, p, n2) == 0) return l.id;
        const int32_t id2 = m->in.intern(p, n2, st);
        const Str &sx = m->in.strs[id2];
        l.p = sx.p; l.n = sx.len; l.id = id2;
        return id2;
    };
    std::vector<int32_t> hier_type_ids;
    int version = o.version;
    std::vector<int32_t> hier_ids;
    auto refresh_hier = [&]() {
        hier_ids.clear();
        hier_type_ids.clear();
        for (auto &h : o.hier) {
            hier_ids.push_back(m->in.intern(h));
            const size_t us = h.find('_');                  // parent_feature.split('_')[0]
            hier_type_ids.push_back(m->in.intern(us == std::string::npos ? h : h.substr(0, us)));
        }
    };
    refresh_hier();
    {   // one row per accepted line at most (+ implicit parents): reserve once instead of doubling 100-byte rows
        int64_t nl_count = 0;
        for (const char *q = T, *e = T + n; q < e; nl_count++) { const char *x = (const char *)memchr(q, '\n', (size_t)(e - q)); if (!x) break; q = x + 1; }
        m->rows.reserve(m->rows.size() + (size_t)nl_count / 2 + 1024);
    }
    int64_t pos = 0;
    while (pos < n && m->status == ST_OK) {
        const char *nl = (const char *)memchr(T + pos, '\n', (size_t)(n - pos));
        const int64_t le = nl ? (nl - T) : n;              // line = [pos, le) (+ the '\n')
        const char *L = T + pos;
        size_t ln = (size_t)(le - pos);
        pos = le + 1;
        if (ln == 0 && !nl) break;
        if (ln > 0 && L[0] == '#') continue;
        if (ln == 0) continue;
        int tabs = 0;
        bool has_cr = false;
        size_t tp[10];
        for (size_t i = 0; i < ln; i++) {
            const char c = L[i];
            if (c == '\t') { if (tabs < 10) tp[tabs] = i; tabs++; }
            else if (c == '\r') has_cr = true;
        }
        if (tabs != 8) continue;
        bool stable = true;
        if (!has_cr && o.replace.empty() && version != 0 &&
            in_list(o.ignore, o.ignore_is_str, L + tp[1] + 1, tp[2] - tp[1] - 1)) continue;       // ignored feature type: nothing else to do
        if (has_cr || !o.replace.empty()) {                // line.replace('\n','').replace('\r','') + features_to_replace, in that order
            clean.assign(L, ln);
            if (has_cr) { tmp.clear(); for (char c : clean) if (c != '\r') tmp.push_back(c); clean.swap(tmp); }
            for (auto &pr : o.replace) {
                if (pr.first.empty()) {                    // str.replace('', x) inserts x between all characters
                    tmp.clear();
                    for (char c : clean) { tmp += pr.second; tmp.push_back(c); }
                    tmp += pr.second;
                    clean.swap(tmp);
                    continue;
                }
                size_t at = 0, hit;
                tmp.clear();
                while ((hit = clean.find(pr.first, at)) != std::string::npos) { tmp.append(clean, at, hit - at); tmp += pr.second; at = hit + pr.first.size(); }
                if (at) { tmp.append(clean, at, std::string::npos); clean.swap(tmp); }
            }
            if (clean.size() != ln || memcmp(clean.data(), L, ln) != 0) {
                m->in.arena.emplace_back(clean);           // the strings of this line point into a private copy
                L = m->in.arena.back().data();
                ln = m->in.arena.back().size();
                stable = true;
            }
        }
        f.clear();
        {
            size_t a = 0;
            for (size_t i = 0; i <= ln; i++) if (i == ln || L[i] == '\t') { f.emplace_back(L + a, i - a); a = i + 1; }
        }
        if (f.size() < 9) { m->status = ST_FEW_FIELDS; break; }
        const char *f8 = f[8].first;
        const size_t f8n = f[8].second;
        if (version == 0) {                                // "auto" (genome.py:288-300)
            if (memchr(f8, '=', f8n)) version = 3;
            else {
                version = 2;
                if (o.has_id_field) {
                    std::string hay = " " + std::string(f8, f8n);
                    for (char &c : hay) if (c == ';') c = ' ';
                    const std::string needle = " " + o.id_field + " ";
                    if (hay.find(needle) == std::string::npos && o.hier.empty()) {
                        o.has_id_field = false;
                        o.has_parent_field = false;
                        const bool g = memmem(f8, f8n, "gene_id", 7) != nullptr, t = memmem(f8, f8n, "transcript_id", 13) != nullptr;
                        if (g && t) { o.hier = {"transcript_id", "gene_id"}; }
                        else if (g) { o.hier = {"gene_id"}; }
                        refresh_hier();
                    }
                }
            }
        }
        if (in_list(o.ignore, o.ignore_is_str, f[2].first, f[2].second)) continue;
        m->n_lines++;
        int64_t c0, c1;
        if (!parse_int(f[3].first, f[3].second, &c0)) { m->status = ST_BAD_INT; m->err_a = m->in.intern(f[3].first, f[3].second, stable); break; }
        if (!parse_int(f[4].first, f[4].second, &c1)) { m->status = ST_BAD_INT; m->err_a = m->in.intern(f[4].first, f[4].second, stable); break; }
        if (c0 > c1) { const int64_t t = c0; c0 = c1; c1 = t; }
        double score = 0;
        bool has_score = false;
        if (!(f[5].second == 1 && f[5].first[0] == '.')) has_score = parse_float(f[5].first, f[5].second, &score);
        int8_t phase = -1;
        if (f[7].second == 1 && f[7].first[0] >= '0' && f[7].first[0] <= '2') phase = (int8_t)(f[7].first[0] - '0');
        // defline_dict (ordered; a repeated key keeps its place and takes the new value)
        kv.clear();
        {
            size_t a = 0;
            for (size_t i = 0; i <= f8n && m->status == ST_OK; i++) {
                if (i != f8n && f8[i] != ';') continue;
                const char *d = f8 + a;
                const size_t dn = i - a;
                a = i + 1;
                if (dn == 0) continue;
                KV e{nullptr, 0, nullptr, 0};
                if (o.has_parent_field && o.parent_field.empty()) { e.k = d; e.kn = 0; e.v = d; e.vn = dn; }
                else if (version == 2) {
                    const char *q = (const char *)memchr(d, '"', dn);
                    size_t x = 0;
                    while (x < dn && is_ws(d[x])) x++;
                    size_t y = x;
                    while (y < dn && !is_ws(d[y])) y++;
                    if (q) {
                        if (x == dn) { m->status = ST_GFF2_NO_KEY; break; }
                        e.k = d + x; e.kn = y - x;
                        const char *q2 = (const char *)memchr(q + 1, '"', dn - (size_t)(q + 1 - d));
                        e.v = q + 1; e.vn = q2 ? (size_t)(q2 - q - 1) : dn - (size_t)(q + 1 - d);
                    } else {
                        size_t x2 = y;
                        while (x2 < dn && is_ws(d[x2])) x2++;
                        size_t y2 = x2;
                        while (y2 < dn && !is_ws(d[y2])) y2++;
                        if (x == dn || x2 == dn) { m->status = ST_GFF2_NO_VALUE; m->err_a = m->in.intern(d, dn, stable); break; }
                        e.k = d + x; e.kn = y - x; e.v = d + x2; e.vn = y2 - x2;
                    }
                } else if (version == 3) {
                    const char *q = (const char *)memchr(d, '=', dn);
                    if (!q) { m->status = ST_GFF3_NO_EQ; break; }
                    e.k = d; e.kn = (size_t)(q - d);
                    const char *q2 = (const char *)memchr(q + 1, '=', dn - (size_t)(q + 1 - d));
                    e.v = q + 1; e.vn = q2 ? (size_t)(q2 - q - 1) : dn - (size_t)(q + 1 - d);
                } else continue;                           // any other gff_version: no attribute is read
                bool found = false;
                for (KV &old : kv) if (old.kn == e.kn && memcmp(old.k, e.k, e.kn) == 0) { old.v = e.v; old.vn = e.vn; found = true; break; }
                if (!found) kv.push_back(e);
            }
            if (m->status != ST_OK) break;
        }
        auto lookup = [&](const std::string &key) -> const KV * {
            for (const KV &e : kv) if (e.kn == key.size() && memcmp(e.k, key.data(), e.kn) == 0) return &e;
            return nullptr;
        };
        int32_t parent = -1;
        if (o.has_parent_field) { if (const KV *e = lookup(o.parent_field)) parent = intern_role(4, e->v, e->vn, stable); }
        else for (auto &h : o.hier) if (const KV *e = lookup(h)) { parent = intern_role(4, e->v, e->vn, stable); break; }
        int32_t id = -1;
        const KV *ide = o.has_id_field ? lookup(o.id_field) : nullptr;
        if (ide) id = m->in.intern(ide->v, ide->vn, stable);
        else if (parent >= 0) {
            const Str ps = m->in.strs[parent];
            scratch.assign(ps.p, ps.len);
            scratch += "-";
            scratch.append(f[2].first, f[2].second);
            id = intern_role(5, scratch.data(), scratch.size(), false);
        } else if (!o.has_id_field) {
            scratch.assign(f[0].first, f[0].second);
            scratch += "-";
            scratch.append(f[2].first, f[2].second);
            scratch.append(f[3].first, f[3].second);
            id = m->in.intern(scratch.data(), scratch.size(), false);
        }
        // (IDfield given, absent from the line and no parent: ID stays None -- the reference then files the object under None;
        //  the Python side refuses that rare case)
        if (id < 0) id = none_id;                          // ID stays None: the object is filed under the key None
        if (m->owner_get(id)) {                            // annotation_set[ID] exists (genome.py:355-364)
            if (id == none_id) { m->status = ST_NONE_TWICE; break; }      // None + '2': TypeError in the reference
            if ((size_t)id >= m->newid_cnt.size()) m->newid_cnt.resize(std::max((size_t)id + 1, m->newid_cnt.size() * 2 + 1024), 0);
            const Str base_id = m->in.strs[id];
            scratch.assign(base_id.p, base_id.len);
            if (m->newid_cnt[id]) { m->newid_cnt[id] += 1; scratch += "-"; scratch += std::to_string(m->newid_cnt[id]); }
            else { m->newid_cnt[id] = 2; scratch += "2"; }
            id = m->in.intern(scratch.data(), scratch.size(), false);
        }
        const int32_t seqid = intern_role(0, f[0].first, f[0].second, stable);
        const int32_t strand = intern_role(1, f[6].first, f[6].second, stable);
        if (parent >= 0) {
            int32_t child_to_assign = id;
            for (size_t hi = 0; hi < o.hier.size(); hi++) {
                const KV *e = lookup(o.hier[hi]);
                if (!e) continue;
                const int32_t pfid = intern_role(6 + (int)(hi & 1), e->v, e->vn, stable);
                const int32_t ptype = hier_type_ids[hi];
                int32_t pparent = -1;
                if (hi + 1 != o.hier.size())
                    for (size_t h2 = hi + 1; h2 < o.hier.size(); h2++) if (const KV *e2 = lookup(o.hier[h2])) pparent = intern_role(6 + (int)(h2 & 1), e2->v, e2->vn, stable);
                const int t = m->table_for(ptype);
                if (t == -2) { m->status = ST_ATTR_CLASH; m->err_a = ptype; break; }
                int64_t *slot = m->filed_get(t, pfid);
                if (slot) {
                    const int64_t prow = m->tables[t].rows[*slot];
                    if (m->rows[prow].is_base) { m->status = ST_PARENT_IS_BASE; m->err_a = pfid; break; }
                    m->add_child(prow, child_to_assign);
                } else {
                    const int64_t prow = m->new_row();
                    Row &pr = m->rows[prow];
                    pr.id = pfid; pr.seqid = seqid; pr.ftype = ptype; pr.strand = strand; pr.parent = pparent; pr.implicit = 1; pr.table = t;
                    m->add_child(prow, child_to_assign);
                    m->file_row(t, pfid, prow);
                    m->reg(pfid, prow);
                }
                child_to_assign = pfid;
            }
            if (m->status != ST_OK) break;
            int64_t *got = m->owner_get(parent);
            if (!got) { m->status = ST_MISSING_PARENT; m->err_a = id; m->err_b = parent; break; }
            if (m->rows[*got].is_base) { m->status = ST_PARENT_IS_BASE; m->err_a = parent; break; }
            m->add_child(*got, id);
        }
        const int32_t ftype = intern_role(2, f[2].first, f[2].second, stable);
        const int t = m->table_for(ftype);
        if (t == -2) { m->status = ST_ATTR_CLASH; m->err_a = ftype; break; }
        const int64_t row = m->new_row();
        Row &r = m->rows[row];
        r.id = id; r.seqid = seqid; r.ftype = ftype; r.strand = strand; r.parent = parent; r.table = t;
        r.source = intern_role(3, f[1].first, f[1].second, stable);
        r.start = c0; r.end = c1; r.score = score; r.has_score = has_score; r.phase = phase;
        r.is_base = in_list(o.base, o.base_is_str, f[2].first, f[2].second) ? 1 : 0;
        for (const KV &e : kv) {
            if (o.has_id_field && e.kn == o.id_field.size() && memcmp(e.k, o.id_field.data(), e.kn) == 0) continue;
            if (o.has_parent_field && e.kn == o.parent_field.size() && memcmp(e.k, o.parent_field.data(), e.kn) == 0) continue;
            const int slot = r.nattr < 2 ? r.nattr : -1;     // the first two attributes of consecutive lines usually repeat
            m->attr_key.push_back(slot >= 0 ? intern_role(8 + slot, e.k, e.kn, stable) : m->in.intern(e.k, e.kn, stable));
            m->attr_val.push_back(slot >= 0 ? intern_role(10 + slot, e.v, e.vn, stable) : m->in.intern(e.v, e.vn, stable));
            r.nattr++;
        }
        m->file_row(t, id, row);
        m->reg(id, row);
    }
    // ---- flat copies for the caller
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    const size_t R = m->rows.size();
    m->c_id.resize(R); m->c_seqid.resize(R); m->c_ftype.resize(R); m->c_strand.resize(R); m->c_source.resize(R); m->c_parent.resize(R);
    m->c_ext.resize(R); m->c_table.resize(R); m->c_nattr.resize(R); m->c_start.resize(R); m->c_end.resize(R); m->c_attr0.resize(R);
    m->c_score.resize(R); m->c_has_score.resize(R); m->c_phase.resize(R); m->c_is_base.resize(R); m->c_implicit.resize(R);
    m->c_child_off.assign(R + 1, 0);
    m->c_ext_children0.resize(R);
    for (size_t i = 0; i < R; i++) {
        const Row &r = m->rows[i];
        m->c_ext_children0[i] = r.ext_children0;
        m->c_id[i] = r.id; m->c_seqid[i] = r.seqid; m->c_ftype[i] = r.ftype; m->c_strand[i] = r.strand; m->c_source[i] = r.source;
        m->c_parent[i] = r.parent; m->c_ext[i] = r.ext; m->c_table[i] = r.table; m->c_nattr[i] = r.nattr; m->c_start[i] = r.start;
        m->c_end[i] = r.end; m->c_attr0[i] = r.attr0; m->c_score[i] = r.score; m->c_has_score[i] = r.has_score; m->c_phase[i] = r.phase;
        m->c_is_base[i] = r.is_base; m->c_implicit[i] = r.implicit;
        m->c_child_off[i + 1] = m->c_child_off[i] + r.n_children;
    }
    m->c_child.reserve((size_t)m->c_child_off[R]);
    for (size_t i = 0; i < R; i++)
        for (int64_t node = m->rows[i].child_head; node >= 0; node = m->kid_next[node]) m->c_child.push_back(m->kid_id[node]);
    m->c_table_off.assign(m->tables.size() + 1, 0);
    for (size_t t = 0; t < m->tables.size(); t++) {
        m->c_table_name.push_back(m->tables[t].name);
        m->c_table_off[t + 1] = m->c_table_off[t] + (int64_t)m->tables[t].rows.size();
        m->c_table_rows.insert(m->c_table_rows.end(), m->tables[t].rows.begin(), m->tables[t].rows.end());
    }
    if (getenv("MG_GFF_TIMING")) { struct timespec ts1; clock_gettime(CLOCK_MONOTONIC, &ts1); fprintf(stderr, "mg_gff_parse: export %.3f s\n", (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec)); }
    *out = m;
    return MG_OK;
}

extern "C" int mg_gff_destroy(mg_gff *m) { delete m; return MG_OK; }

// info[0] status, [1] err_a, [2] err_b, [3] rows, [4] strings, [5] attributes, [6] child entries, [7] tables, [8] accepted lines,
// [9] IDfield still set, [10] parent_field still set, [11] the string id that stands for an ID of None
extern "C" int mg_gff_info(mg_gff *m, int64_t *info) {
    if (!m || !info) { mg_set_error("mg_gff_info: NULL argument"); return MG_EINVAL; }
    info[0] = m->status; info[1] = m->err_a; info[2] = m->err_b; info[3] = (int64_t)m->rows.size(); info[4] = (int64_t)m->in.strs.size();
    info[5] = (int64_t)m->attr_key.size(); info[6] = (int64_t)m->c_child.size(); info[7] = (int64_t)m->tables.size(); info[8] = m->n_lines;
    info[9] = m->o.has_id_field; info[10] = m->o.has_parent_field; info[11] = m->none_id;
    return MG_OK;
}

// column by name: pointer to the library-owned array, element count and element size (valid until mg_gff_destroy)
extern "C" int mg_gff_column(mg_gff *m, const char *name, const void **ptr, int64_t *n, int32_t *elem) {
    if (!m || !name || !ptr || !n || !elem) { mg_set_error("mg_gff_column: NULL argument"); return MG_EINVAL; }
#define COL(nm, v) if (!strcmp(name, nm)) { *ptr = (v).data(); *n = (int64_t)(v).size(); *elem = (int32_t)sizeof((v)[0]); return MG_OK; }
    COL("id", m->c_id) COL("seqid", m->c_seqid) COL("ftype", m->c_ftype) COL("strand", m->c_strand) COL("source", m->c_source)
    COL("parent", m->c_parent) COL("ext", m->c_ext) COL("table", m->c_table) COL("nattr", m->c_nattr) COL("start", m->c_start)
    COL("end", m->c_end) COL("attr0", m->c_attr0) COL("score", m->c_score) COL("has_score", m->c_has_score) COL("phase", m->c_phase)
    COL("ext_children0", m->c_ext_children0) COL("is_base", m->c_is_base) COL("implicit", m->c_implicit) COL("child_off", m->c_child_off) COL("child", m->c_child)
    COL("attr_key", m->attr_key) COL("attr_val", m->attr_val) COL("table_name", m->c_table_name) COL("table_off", m->c_table_off)
    COL("table_rows", m->c_table_rows)
#undef COL
    mg_set_error("mg_gff_column: unknown column '%s'", name);
    return MG_EINVAL;
}

// strings ids[0..n) copied back to back into `pool` (capacity cap); off[n+1] receives their offsets.  pool == NULL: sizes only.
extern "C" int mg_gff_strings(mg_gff *m, const int32_t *ids, int64_t n, uint8_t *pool, int64_t cap, int64_t *off) {
    if (!m || (!ids && n > 0) || !off) { mg_set_error("mg_gff_strings: NULL argument"); return MG_EINVAL; }
    int64_t o = 0;
    for (int64_t i = 0; i < n; i++) {
        off[i] = o;
        if (ids[i] < 0 || (size_t)ids[i] >= m->in.strs.size()) { mg_set_error("mg_gff_strings: bad string id"); return MG_EINVAL; }
        const Str &s = m->in.strs[ids[i]];
        if (pool) { if (o + s.len > cap) { mg_set_error("mg_gff_strings: pool too small"); return MG_EINVAL; } memcpy(pool + o, s.p, s.len); }
        o += s.len;
    }
    off[n] = o;
    return MG_OK;
}

// id of a string the model already holds (-1: unknown)
extern "C" int64_t mg_gff_find(mg_gff *m, const uint8_t *s, int64_t n) {
    if (!m || (!s && n > 0)) return -1;
    return m->in.find((const char *)s, (size_t)n, fnv1a((const char *)s, (size_t)n));
}

// ---- iteration order of a CPython-2.7 dict (the reference prints records by iterating plain dicts: genome.py:580) -----------
// Replays Objects/dictobject.c of CPython 2.7 on the keys' 2.7 string hashes (hash randomisation off): 8 initial slots, probe
// i = 5 i + perturb + 1 with perturb >>= 5, grow when fill * 3 >= 2 * size to the first power of two > 4 * used (2 * used above
// 50 000 entries).  rounds = 2 re-inserts the keys in the first table's slot order, which is what read_gff's copy.deepcopy does
// (genome.py:415).  perm[k] = index (into the input) of the k-th key the dict yields.  Host only.
static inline uint64_t py27_string_hash(const unsigned char *p, size_t len) {
    if (len == 0) return 0;
    uint64_t x = (uint64_t)p[0] << 7;
    for (size_t i = 0; i < len; i++) x = (1000003ull * x) ^ p[i];
    x ^= (uint64_t)len;
    if (x == ~0ull) x = ~0ull - 1;                    // -1 is CPython's error value: -2
    return x;
}

static inline void py27_place(std::vector<int64_t> &tab, size_t mask, const std::vector<uint64_t> &h, int64_t k) {
    const uint64_t hh = h[k];
    size_t i = hh & mask;
    if (tab[i] != -1) {
        uint64_t perturb = hh, j = i;
        while (true) {
            j = (j << 2) + j + perturb + 1;
            perturb >>= 5;
            i = j & mask;
            if (tab[i] == -1) break;
        }
    }
    tab[i] = k;
}

static void py27_dict_order(const std::vector<uint64_t> &h, std::vector<int64_t> &order) {     // order: insertion order in, iteration order out
    size_t size = 8, used = 0;
    std::vector<int64_t> slots(size, -1), grown;
    for (int64_t k : order) {
        py27_place(slots, size - 1, h, k);
        used++;
        if (used * 3 >= size * 2) {
            const size_t minused = (used > 50000 ? 2 : 4) * used;
            size_t newsize = 8;
            while (newsize <= minused) newsize <<= 1;
            grown.assign(newsize, -1);
            for (int64_t k2 : slots) if (k2 != -1) py27_place(grown, newsize - 1, h, k2);
            slots.swap(grown);
            size = newsize;
        }
    }
    size_t o = 0;
    for (int64_t k : slots) if (k != -1) order[o++] = k;
}

extern "C" int mg_py2_order(const uint8_t *pool, const int64_t *off, int64_t n, int rounds, int64_t *perm) {
    if (n < 0 || (n > 0 && (!off || !perm)) || rounds < 1) { mg_set_error("mg_py2_order: bad argument"); return MG_EINVAL; }
    std::vector<uint64_t> h((size_t)n);
    for (int64_t i = 0; i < n; i++) h[i] = py27_string_hash(pool + off[i], (size_t)(off[i + 1] - off[i]));
    std::vector<int64_t> order((size_t)n);
    for (int64_t i = 0; i < n; i++) order[i] = i;
    for (int r = 0; r < rounds; r++) py27_dict_order(h, order);
    for (int64_t i = 0; i < n; i++) perm[i] = order[i];
    return MG_OK;
}

// the same on string ids of a model (no Python string is built)
extern "C" int mg_gff_py2_order(mg_gff *m, const int32_t *ids, int64_t n, int rounds, int64_t *perm) {
    if (!m || n < 0 || (n > 0 && (!ids || !perm)) || rounds < 1) { mg_set_error("mg_gff_py2_order: bad argument"); return MG_EINVAL; }
    std::vector<uint64_t> h((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        if (ids[i] < 0 || (size_t)ids[i] >= m->in.strs.size()) { mg_set_error("mg_gff_py2_order: bad string id"); return MG_EINVAL; }
        const Str &t = m->in.strs[ids[i]];
        h[i] = py27_string_hash((const unsigned char *)t.p, t.len);
    }
    std::vector<int64_t> order((size_t)n);
    for (int64_t i = 0; i < n; i++) order[i] = i;
    for (int r = 0; r < rounds; r++) py27_dict_order(h, order);
    for (int64_t i = 0; i < n; i++) perm[i] = order[i];
    return MG_OK;
}

// ---- flattener: AnnotationSet.get_fasta(feature) on the integer model ------------------------------------------------------
// tops[0..n_top): rows of the feature table in the set's iteration order.  contig_of[string id] = contig index of that seqid in
// the packed genome, -1 = not a contig.  Emits, per top, the records ParentAnnotation.get_fasta would (genome.py:683-719):
// base children keyed by coords (identical coords collapse, last wins), sorted ascending, reversed when the LAST child's strand
// is '-', each segment reverse-complemented by its OWN strand; parent children recursed; a top without records -> a blank line.
// framing != 0 builds the literals '>' + ID + '\n' ... '\n' (genome.py:710, :727, :581).
// status (info[0]) != 0: the model needs the object path (mixed children, unknown child, unknown seqid, odd strand):
//   1 = unknown child ID, 2 = mixed base/parent children, 3 = strand not in "+.-", 4 = seqid is not a contig
struct mg_gff_flat {
    std::vector<int64_t> rec_seg_off{0}, seg_start, seg_end, rec_lit_off, top_rec_off{0}, entry_off;
    std::vector<int32_t> seg_contig, rec_pre, rec_suf, rec_name;
    std::vector<int8_t> seg_strand, rec_phase;
    std::vector<uint8_t> lit;
    int status = 0;
    int64_t err = -1;
    int64_t st[2] = {0, -1};
};

namespace {
struct SegTmp { int64_t s, e; int64_t row; int64_t ord; };

static void collect(mg_gff *m, mg_gff_flat *fl, int64_t row, const int32_t *contig_of, int64_t n_contig_of, std::vector<SegTmp> &tmp, int depth) {
    const Row &r = m->rows[row];
    if (r.n_children == 0 || fl->status || depth > 64) return;
    const int32_t first_child = m->kid_id[r.child_head];
    int64_t *f0 = m->owner_get(first_child);
    if (!f0) { fl->status = 1; fl->err = first_child; return; }
    if (m->rows[*f0].is_base) {
        tmp.clear();
        int32_t last_strand = -1;
        for (int64_t node = r.child_head; node >= 0; node = m->kid_next[node]) {
            const int32_t c = m->kid_id[node];
            int64_t *cr = m->owner_get(c);
            if (!cr) { fl->status = 1; fl->err = c; return; }
            const Row &ch = m->rows[*cr];
            if (!ch.is_base) { fl->status = 2; fl->err = r.id; return; }
            tmp.push_back({ch.start, ch.end, *cr, (int64_t)tmp.size()});
            last_strand = ch.strand;
        }
        // child_dict[coords] = child: identical coords collapse (the last one wins); sorted(child_dict): ascending (start, end)
        std::stable_sort(tmp.begin(), tmp.end(), [](const SegTmp &a, const SegTmp &b) { return a.s < b.s || (a.s == b.s && a.e < b.e); });
        {
            size_t w = 0;
            for (size_t i = 0; i < tmp.size(); i++) {
                if (i + 1 < tmp.size() && tmp[i + 1].s == tmp[i].s && tmp[i + 1].e == tmp[i].e) continue;   // a later twin follows
                tmp[w++] = tmp[i];
            }
            tmp.resize(w);
        }
        const Str &ls = m->in.strs[last_strand];
        const bool rev = ls.len == 1 && ls.p[0] == '-';
        int8_t phase = 0;
        for (size_t q = 0; q < tmp.size(); q++) {
            const SegTmp &t = tmp[rev ? tmp.size() - 1 - q : q];
            const Row &ch = m->rows[t.row];
            const Str &ss = m->in.strs[ch.strand];
            int8_t minus;
            if (ss.len == 1 && (ss.p[0] == '+' || ss.p[0] == '.')) minus = 0;
            else if (ss.len == 1 && ss.p[0] == '-') minus = 1;
            else { fl->status = 3; fl->err = ch.id; return; }
            const int32_t cg = (ch.seqid >= 0 && ch.seqid < n_contig_of) ? contig_of[ch.seqid] : -1;
            if (cg < 0) { fl->status = 4; fl->err = ch.seqid; return; }
            fl->seg_contig.push_back(cg);
            fl->seg_start.push_back(t.s);
            fl->seg_end.push_back(t.e);
            fl->seg_strand.push_back(minus);
            if (q == 0) phase = ch.phase > 0 ? ch.phase : 0;
        }
        fl->rec_seg_off.push_back((int64_t)fl->seg_contig.size());
        fl->rec_phase.push_back(phase);
        fl->rec_name.push_back(r.id);
        return;
    }
    for (int64_t node = r.child_head; node >= 0; node = m->kid_next[node]) {
        const int32_t c = m->kid_id[node];
        int64_t *cr = m->owner_get(c);
        if (!cr) { fl->status = 1; fl->err = c; return; }
        if (m->rows[*cr].is_base) { fl->status = 2; fl->err = r.id; return; }
        collect(m, fl, *cr, contig_of, n_contig_of, tmp, depth + 1);
        if (fl->status) return;
    }
}
}  // namespace

extern "C" int mg_gff_flatten(mg_gff *m, const int64_t *tops, int64_t n_top, const int32_t *contig_of, int64_t n_contig_of,
                              int framing, mg_gff_flat **out) {
    if (!m || (!tops && n_top > 0) || !out) { mg_set_error("mg_gff_flatten: NULL argument"); return MG_EINVAL; }
    mg_gff_flat *fl = new mg_gff_flat();
    std::vector<SegTmp> tmp;
    for (int64_t k = 0; k < n_top && !fl->status; k++) {
        if (tops[k] < 0 || (size_t)tops[k] >= m->rows.size()) { delete fl; mg_set_error("mg_gff_flatten: bad row"); return MG_EINVAL; }
        const size_t before = fl->rec_name.size();
        if (!m->rows[tops[k]].is_base) collect(m, fl, tops[k], contig_of, n_contig_of, tmp, 0);
        else { fl->status = 2; fl->err = m->rows[tops[k]].id; }
        if (fl->rec_name.size() == before) {               // "" in the joined list: a blank line
            fl->rec_seg_off.push_back((int64_t)fl->seg_contig.size());
            fl->rec_phase.push_back(0);
            fl->rec_name.push_back(-1);
        }
        fl->top_rec_off.push_back((int64_t)fl->rec_name.size());
    }
    const size_t R = fl->rec_name.size();
    fl->rec_lit_off.resize(R); fl->rec_pre.resize(R); fl->rec_suf.resize(R);
    for (size_t r = 0; r < R; r++) {
        fl->rec_lit_off[r] = (int64_t)fl->lit.size();
        if (!framing) { fl->rec_pre[r] = fl->rec_suf[r] = 0; continue; }
        if (fl->rec_name[r] < 0) { fl->rec_pre[r] = 0; fl->rec_suf[r] = 1; fl->lit.push_back('\n'); continue; }
        const Str &s = m->in.strs[fl->rec_name[r]];
        fl->lit.push_back('>');
        fl->lit.insert(fl->lit.end(), (const uint8_t *)s.p, (const uint8_t *)s.p + s.len);
        fl->lit.push_back('\n');
        fl->lit.push_back('\n');
        fl->rec_pre[r] = (int32_t)s.len + 2;
        fl->rec_suf[r] = 1;
    }
    *out = fl;
    return MG_OK;
}

extern "C" int mg_gff_flat_destroy(mg_gff_flat *f) { delete f; return MG_OK; }

extern "C" int mg_gff_flat_column(mg_gff_flat *f, const char *name, const void **ptr, int64_t *n, int32_t *elem) {
    if (!f || !name || !ptr || !n || !elem) { mg_set_error("mg_gff_flat_column: NULL argument"); return MG_EINVAL; }
    if (!strcmp(name, "status")) { f->st[0] = f->status; f->st[1] = f->err; *ptr = f->st; *n = 2; *elem = 8; return MG_OK; }
#define COL(nm, v) if (!strcmp(name, nm)) { *ptr = (v).data(); *n = (int64_t)(v).size(); *elem = (int32_t)sizeof((v)[0]); return MG_OK; }
    COL("rec_seg_off", f->rec_seg_off) COL("seg_contig", f->seg_contig) COL("seg_start", f->seg_start) COL("seg_end", f->seg_end)
    COL("seg_strand", f->seg_strand) COL("rec_phase", f->rec_phase) COL("rec_name", f->rec_name) COL("rec_lit_off", f->rec_lit_off)
    COL("rec_pre", f->rec_pre) COL("rec_suf", f->rec_suf) COL("lit", f->lit) COL("top_rec_off", f->top_rec_off)
#undef COL
    mg_set_error("mg_gff_flat_column: unknown column '%s'", name);
    return MG_EINVAL;
}
