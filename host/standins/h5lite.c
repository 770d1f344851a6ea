/* h5lite.c -- a native writer of HDF5 files behind the subset of the HDF5 C API declared in include/hdf5.h, so that
 * the reference's own hdf5_funcs.c (file creation hdf5_funcs.c:33-326, WriteDataToFile :492-779, end-of-run series
 * :1028-1225, compound {r,i} type :1302-1333) runs unmodified where libhdf5 does not exist.
 *
 * File format written (HDF5 File Format Specification, version 0 structures, readable by every libhdf5 / h5py):
 *   superblock v0 (8-byte offsets and lengths) -> root symbol-table entry -> version-1 object headers;
 *   groups   = symbol table message -> v1 B-tree ("TREE") -> symbol nodes ("SNOD") + local heap ("HEAP") of names;
 *   datasets = dataspace v1 + datatype v1 (IEEE f64 LE, i32 LE, compound) + fill value v2 + contiguous layout v3;
 *   attributes = attribute message v1 in the owner's object header.
 * Raw data are written where they are produced; all metadata are kept in memory and written as one block behind the
 * raw data on H5Fclose (the superblock then points at the new root).  A file created by this process can be reopened
 * with H5Fopen(H5F_ACC_RDWR) by the same process, which is what the reference does on every save (hdf5_funcs.c:538);
 * the raw data of the new save then overwrite the superseded metadata block, so nothing is wasted.
 * Single process, hyperslab selections with start/count only (all the reference uses), no chunking, no deletion.
 *
 * Read side (restart from a saved state, host/nsb200_hooks.c): H5Fopen(H5F_ACC_RDONLY) of a file this process did not
 * create parses the same structures back into the in-memory model (anything else in the file is refused), after which
 * H5Lexists / H5Gopen / H5Dopen / H5Dget_space / H5Dread work on it; such a file is never written to. */
#include "hdf5.h"
#include "hdf5_hl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LEAF_K 16u      /* symbol nodes hold up to 2 * LEAF_K entries */
#define INT_K 16u       /* B-tree nodes hold up to 2 * INT_K children */
#define UNDEF 0xffffffffffffffffull
#define MAXRANK 8

/* ------------------------------------------------------------------------------------------------ model */
typedef struct { int kind; /* 0 i32, 1 f64, 6 compound */ size_t size; int nmem; char mname[8][32]; size_t moff[8]; int mkind[8]; } h5t;
typedef struct { char name[64]; h5t type; int rank; hsize_t dims[MAXRANK]; void* data; size_t bytes; } h5attr;
typedef struct h5node {
    char* name;
    int is_group;
    struct h5node** child; int nchild, capchild;
    h5attr* attr; int nattr;
    h5t type; int rank; hsize_t dims[MAXRANK]; uint64_t addr, bytes;   /* dataset */
    uint64_t ohdr, btree, heap;                                         /* filled while serialising */
} h5node;
typedef struct { char path[2048]; FILE* fp; h5node* root; uint64_t data_end; int open; int readonly; } h5file;
typedef struct { int rank; hsize_t dims[MAXRANK]; int sel; hsize_t start[MAXRANK], count[MAXRANK]; } h5space;

enum { K_FREE = 0, K_FILE, K_GROUP, K_DSET, K_SPACE, K_TYPE, K_ATTR, K_PLIST };
typedef struct { int kind; void* ptr; h5file* file; h5node* owner; } h5obj;
static h5obj g_obj[4096];
static h5file* g_files[64];

static hid_t new_id(int kind, void* ptr, h5file* f, h5node* owner) {
    for (int i = 0; i < 4096; ++i)
        if (g_obj[i].kind == K_FREE) { g_obj[i].kind = kind; g_obj[i].ptr = ptr; g_obj[i].file = f; g_obj[i].owner = owner; return 100 + i; }
    return -1;
}
static h5obj* get(hid_t id, int kind) {
    if (id < 100 || id >= 100 + 4096) return NULL;
    h5obj* o = &g_obj[id - 100];
    return (o->kind == kind) ? o : NULL;
}
static void drop(hid_t id) { if (id >= 100 && id < 100 + 4096) g_obj[id - 100].kind = K_FREE; }

static int resolve_type(hid_t id, h5t* out) {
    memset(out, 0, sizeof *out);
    if (id == H5T_NATIVE_DOUBLE) { out->kind = 1; out->size = 8; return 0; }
    if (id == H5T_NATIVE_INT) { out->kind = 0; out->size = 4; return 0; }
    h5obj* o = get(id, K_TYPE);
    if (!o) return -1;
    *out = *(h5t*)o->ptr;
    return 0;
}

/* location id (file or group) -> node */
static h5node* loc_node(hid_t loc, h5file** f) {
    h5obj* o = get(loc, K_FILE);
    if (o) { *f = (h5file*)o->ptr; return (*f)->root; }
    o = get(loc, K_GROUP);
    if (o) { *f = o->file; return (h5node*)o->ptr; }
    return NULL;
}
static h5node* find_child(h5node* g, const char* name, size_t len) {
    for (int i = 0; i < g->nchild; ++i)
        if (strlen(g->child[i]->name) == len && !strncmp(g->child[i]->name, name, len)) return g->child[i];
    return NULL;
}
/* walks "a/b/c" below `from` ("/..." starts at the root); returns the parent of the last component and its name */
static h5node* walk(h5file* f, h5node* from, const char* path, const char** last, size_t* lastlen) {
    h5node* cur = (path[0] == '/') ? f->root : from;
    while (*path == '/') ++path;
    for (;;) {
        const char* s = strchr(path, '/');
        if (!s) { *last = path; *lastlen = strlen(path); return cur; }
        h5node* nx = find_child(cur, path, (size_t)(s - path));
        if (!nx || !nx->is_group) return NULL;
        cur = nx;
        path = s + 1;
        while (*path == '/') ++path;
    }
}
static h5node* add_child(h5node* g, const char* name, size_t len, int is_group) {
    if (g->nchild == g->capchild) {
        g->capchild = g->capchild ? 2 * g->capchild : 8;
        g->child = (h5node**)realloc(g->child, sizeof(h5node*) * (size_t)g->capchild);
    }
    h5node* n = (h5node*)calloc(1, sizeof *n);
    n->name = (char*)malloc(len + 1);
    memcpy(n->name, name, len);
    n->name[len] = 0;
    n->is_group = is_group;
    g->child[g->nchild++] = n;
    return n;
}
static void free_node(h5node* n) {
    for (int i = 0; i < n->nchild; ++i) free_node(n->child[i]);
    for (int i = 0; i < n->nattr; ++i) free(n->attr[i].data);
    free(n->child); free(n->attr); free(n->name); free(n);
}

/* ------------------------------------------------------------------------------------------------ byte buffer */
typedef struct { unsigned char* p; size_t n, cap; uint64_t base; } blob;
static void b_need(blob* b, size_t more) {
    if (b->n + more > b->cap) { b->cap = (b->n + more) * 2 + 4096; b->p = (unsigned char*)realloc(b->p, b->cap); }
}
static void b_put(blob* b, const void* src, size_t n) { b_need(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void b_zero(blob* b, size_t n) { b_need(b, n); memset(b->p + b->n, 0, n); b->n += n; }
static void b_u8(blob* b, unsigned v) { unsigned char c = (unsigned char)v; b_put(b, &c, 1); }
static void b_u16(blob* b, unsigned v) { unsigned char c[2] = {(unsigned char)v, (unsigned char)(v >> 8)}; b_put(b, c, 2); }
static void b_u32(blob* b, uint32_t v) { unsigned char c[4]; for (int i = 0; i < 4; ++i) c[i] = (unsigned char)(v >> (8 * i)); b_put(b, c, 4); }
static void b_u64(blob* b, uint64_t v) { unsigned char c[8]; for (int i = 0; i < 8; ++i) c[i] = (unsigned char)(v >> (8 * i)); b_put(b, c, 8); }
static void b_align8(blob* b) { while (b->n % 8) b_u8(b, 0); }
static uint64_t b_addr(const blob* b) { return b->base + b->n; }

/* ------------------------------------------------------------------------------------------------ messages */
static void put_datatype(blob* b, const h5t* t) {
    if (t->kind == 1) {                       /* IEEE 754 binary64, little endian (class 1, version 1) */
        b_u8(b, 0x11); b_u8(b, 0x20); b_u8(b, 63); b_u8(b, 0);   /* LE, no padding, mantissa msb implied, sign at bit 63 */
        b_u32(b, 8);
        b_u16(b, 0); b_u16(b, 64); b_u8(b, 52); b_u8(b, 11); b_u8(b, 0); b_u8(b, 52); b_u32(b, 1023);
    } else if (t->kind == 0) {                /* two's complement int32, little endian (class 0, version 1) */
        b_u8(b, 0x10); b_u8(b, 0x08); b_u8(b, 0); b_u8(b, 0);
        b_u32(b, 4);
        b_u16(b, 0); b_u16(b, 32);
    } else {                                  /* compound (class 6, version 1) */
        b_u8(b, 0x16); b_u8(b, (unsigned)t->nmem & 0xff); b_u8(b, ((unsigned)t->nmem >> 8) & 0xff); b_u8(b, 0);
        b_u32(b, (uint32_t)t->size);
        for (int m = 0; m < t->nmem; ++m) {
            const size_t ln = strlen(t->mname[m]) + 1;
            b_put(b, t->mname[m], ln);
            b_zero(b, (8 - ln % 8) % 8);       /* name padded to a multiple of 8 */
            b_u32(b, (uint32_t)t->moff[m]);
            b_u8(b, 0); b_zero(b, 3);          /* dimensionality 0 */
            b_u32(b, 0); b_u32(b, 0);          /* dimension permutation, reserved */
            b_zero(b, 16);                     /* four dimension sizes */
            h5t mt; memset(&mt, 0, sizeof mt); mt.kind = t->mkind[m]; mt.size = mt.kind ? 8 : 4;
            put_datatype(b, &mt);
        }
    }
}
static void put_dataspace(blob* b, int rank, const hsize_t* dims) {
    b_u8(b, 1); b_u8(b, (unsigned)rank); b_u8(b, 0); b_u8(b, 0); b_u32(b, 0);   /* version 1, no max dims */
    for (int d = 0; d < rank; ++d) b_u64(b, dims[d]);
}
/* one version-1 header message: type, data built by `fill` into a scratch blob, padded to 8 bytes */
static void put_message(blob* out, unsigned type, const blob* data) {
    const size_t padded = (data->n + 7) / 8 * 8;
    b_u16(out, type); b_u16(out, (unsigned)padded); b_u8(out, 0); b_zero(out, 3);
    b_put(out, data->p, data->n);
    b_zero(out, padded - data->n);
}
static void put_attr_messages(blob* msgs, const h5node* n, int* count) {
    for (int i = 0; i < n->nattr; ++i) {
        const h5attr* a = &n->attr[i];
        blob dt = {0}, ds = {0}, m = {0};
        put_datatype(&dt, &a->type);
        put_dataspace(&ds, a->rank, a->dims);
        const size_t ln = strlen(a->name) + 1;
        b_u8(&m, 1); b_u8(&m, 0); b_u16(&m, (unsigned)ln); b_u16(&m, (unsigned)dt.n); b_u16(&m, (unsigned)ds.n);
        b_put(&m, a->name, ln); b_align8(&m);
        b_put(&m, dt.p, dt.n); b_align8(&m);
        b_put(&m, ds.p, ds.n); b_align8(&m);
        b_put(&m, a->data, a->bytes);
        put_message(msgs, 0x000c, &m);
        ++*count;
        free(dt.p); free(ds.p); free(m.p);
    }
}
static uint64_t put_object_header(blob* b, const blob* msgs, int nmsg) {
    b_align8(b);
    const uint64_t at = b_addr(b);
    b_u8(b, 1); b_u8(b, 0); b_u16(b, (unsigned)nmsg); b_u32(b, 1); b_u32(b, (uint32_t)msgs->n); b_u32(b, 0);   /* 12 bytes + pad to 16 */
    b_put(b, msgs->p, msgs->n);
    return at;
}

/* ------------------------------------------------------------------------------------------------ groups */
static int cmp_nodes(const void* x, const void* y) { return strcmp((*(h5node* const*)x)->name, (*(h5node* const*)y)->name); }

static void put_symbol_entry(blob* b, uint64_t name_off, const h5node* c) {
    b_u64(b, name_off); b_u64(b, c->ohdr);
    if (c->is_group) { b_u32(b, 1); b_u32(b, 0); b_u64(b, c->btree); b_u64(b, c->heap); }
    else { b_u32(b, 0); b_u32(b, 0); b_zero(b, 16); }
}

static void serialize(blob* b, h5node* n) {
    if (!n->is_group) {
        blob msgs = {0}, m = {0};
        int cnt = 0;
        put_dataspace(&m, n->rank, n->dims); put_message(&msgs, 0x0001, &m); ++cnt; m.n = 0;
        put_datatype(&m, &n->type); put_message(&msgs, 0x0003, &m); ++cnt; m.n = 0;
        b_u8(&m, 2); b_u8(&m, 2); b_u8(&m, 2); b_u8(&m, 1); b_u32(&m, 0);            /* fill value v2: late alloc, write if set, default value */
        put_message(&msgs, 0x0005, &m); ++cnt; m.n = 0;
        b_u8(&m, 3); b_u8(&m, 1); b_u64(&m, n->bytes ? n->addr : UNDEF); b_u64(&m, n->bytes);   /* layout v3, contiguous */
        put_message(&msgs, 0x0008, &m); ++cnt;
        put_attr_messages(&msgs, n, &cnt);
        n->ohdr = put_object_header(b, &msgs, cnt);
        free(msgs.p); free(m.p);
        return;
    }
    for (int i = 0; i < n->nchild; ++i) serialize(b, n->child[i]);
    qsort(n->child, (size_t)n->nchild, sizeof(h5node*), cmp_nodes);
    /* local heap: offset 0 holds the empty name */
    blob names = {0};
    b_zero(&names, 8);
    uint64_t* noff = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n->nchild + 1));
    for (int i = 0; i < n->nchild; ++i) {
        noff[i] = names.n;
        b_put(&names, n->child[i]->name, strlen(n->child[i]->name) + 1);
        b_align8(&names);
    }
    b_align8(b);
    n->heap = b_addr(b);
    b_put(b, "HEAP", 4); b_u8(b, 0); b_zero(b, 3);
    b_u64(b, names.n); b_u64(b, 1 /* H5HL_FREE_NULL: no free block */); b_u64(b, n->heap + 32);
    b_put(b, names.p, names.n);
    free(names.p);
    /* symbol nodes */
    const int per = 2 * (int)LEAF_K;
    int nsnod = (n->nchild + per - 1) / per;
    if (nsnod == 0) nsnod = 1;
    uint64_t* addr = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nsnod);
    uint64_t* lastkey = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nsnod);   /* heap offset of the largest name below each child */
    for (int s = 0; s < nsnod; ++s) {
        const int lo = s * per, hi = (lo + per < n->nchild) ? lo + per : n->nchild;
        b_align8(b);
        addr[s] = b_addr(b);
        b_put(b, "SNOD", 4); b_u8(b, 1); b_u8(b, 0); b_u16(b, (unsigned)(hi - lo));
        for (int i = lo; i < hi; ++i) put_symbol_entry(b, noff[i], n->child[i]);
        b_zero(b, (size_t)(per - (hi - lo)) * 40);
        lastkey[s] = (hi > lo) ? noff[hi - 1] : 0;
    }
    /* B-tree levels, bottom up */
    int count = nsnod, level = 0;
    for (;;) {
        const int fan = 2 * (int)INT_K;
        const int nnode = (count + fan - 1) / fan;
        uint64_t* naddr = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nnode);
        uint64_t* nlast = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nnode);
        const size_t node_bytes = 24 + (size_t)(2 * fan + 1) * 8;
        b_align8(b);
        const uint64_t first = b_addr(b);
        for (int k = 0; k < nnode; ++k) {
            const int lo = k * fan, hi = (lo + fan < count) ? lo + fan : count;
            naddr[k] = first + (uint64_t)k * node_bytes;
            b_put(b, "TREE", 4); b_u8(b, 0); b_u8(b, (unsigned)level); b_u16(b, (unsigned)(hi - lo));
            b_u64(b, k > 0 ? naddr[k - 1] : UNDEF);
            b_u64(b, k + 1 < nnode ? first + (uint64_t)(k + 1) * node_bytes : UNDEF);
            b_u64(b, lo > 0 ? lastkey[lo - 1] : 0);            /* key 0: everything in child 0 is greater than this name */
            for (int c = lo; c < hi; ++c) { b_u64(b, addr[c]); b_u64(b, lastkey[c]); }
            b_zero(b, (size_t)(fan - (hi - lo)) * 16);
            nlast[k] = lastkey[hi - 1];
        }
        free(addr); free(lastkey);
        addr = naddr; lastkey = nlast; count = nnode; ++level;
        if (count == 1) break;
    }
    n->btree = addr[0];
    free(addr); free(lastkey); free(noff);
    blob msgs = {0}, m = {0};
    int cnt = 0;
    b_u64(&m, n->btree); b_u64(&m, n->heap);
    put_message(&msgs, 0x0011, &m); ++cnt;
    put_attr_messages(&msgs, n, &cnt);
    n->ohdr = put_object_header(b, &msgs, cnt);
    free(msgs.p); free(m.p);
}

static int flush_metadata(h5file* f) {
    blob b = {0};
    b.base = (f->data_end + 7) / 8 * 8;
    serialize(&b, f->root);
    b_align8(&b);
    const uint64_t eof = b.base + b.n;
    blob sb = {0};
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    b_put(&sb, sig, 8);
    b_u8(&sb, 0); b_u8(&sb, 0); b_u8(&sb, 0); b_u8(&sb, 0); b_u8(&sb, 0);       /* superblock, free space, root entry, reserved, shared header versions */
    b_u8(&sb, 8); b_u8(&sb, 8); b_u8(&sb, 0);                                    /* sizes of offsets and lengths */
    b_u16(&sb, LEAF_K); b_u16(&sb, INT_K); b_u32(&sb, 0);
    b_u64(&sb, 0); b_u64(&sb, UNDEF); b_u64(&sb, eof); b_u64(&sb, UNDEF);        /* base, free space info, end of file, driver info */
    b_u64(&sb, 0); b_u64(&sb, f->root->ohdr); b_u32(&sb, 1); b_u32(&sb, 0); b_u64(&sb, f->root->btree); b_u64(&sb, f->root->heap);
    int ok = (fseek(f->fp, (long)b.base, SEEK_SET) == 0) && (fwrite(b.p, 1, b.n, f->fp) == b.n) &&
             (fseek(f->fp, 0, SEEK_SET) == 0) && (fwrite(sb.p, 1, sb.n, f->fp) == sb.n) && (fflush(f->fp) == 0);
    free(b.p); free(sb.p);
    return ok ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------ loader (read side) */
typedef struct { FILE* fp; uint64_t eof; unsigned leaf_k, int_k; int depth; } h5rd;
static uint64_t le(const unsigned char* p, int n) { uint64_t v = 0; for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i]; return v; }
static int rd_at(h5rd* r, uint64_t at, void* dst, size_t n) {
    if (at > r->eof || n > r->eof - at) return -1;
    return (fseek(r->fp, (long)at, SEEK_SET) == 0 && fread(dst, 1, n, r->fp) == n) ? 0 : -1;
}
/* datatype message v1 at p (avail bytes): IEEE f64 LE, plain little-endian i32, compound of those; *used = encoded size */
static int rd_datatype(const unsigned char* p, size_t avail, h5t* t, size_t* used) {
    memset(t, 0, sizeof *t);
    if (avail < 8) return -1;
    const unsigned cls = p[0] & 15u, ver = p[0] >> 4;
    const uint64_t size = le(p + 4, 4);
    if (ver != 1) return -1;
    if (cls == 0) {
        if (avail < 12 || size != 4 || (p[1] & 1u) || !(p[1] & 8u) || le(p + 8, 2) != 0 || le(p + 10, 2) != 32) return -1;
        t->kind = 0; t->size = 4; *used = 12;
        return 0;
    }
    if (cls == 1) {
        if (avail < 20 || size != 8 || p[1] != 0x20 || p[2] != 63 || le(p + 8, 2) != 0 || le(p + 10, 2) != 64 || p[12] != 52 || p[13] != 11 ||
            p[14] != 0 || p[15] != 52 || le(p + 16, 4) != 1023) return -1;
        t->kind = 1; t->size = 8; *used = 20;
        return 0;
    }
    if (cls == 6) {
        const unsigned nmem = p[1] | ((unsigned)p[2] << 8);
        if (nmem < 1 || nmem > 8) return -1;
        t->kind = 6; t->size = (size_t)size; t->nmem = (int)nmem;
        size_t q = 8;
        for (unsigned m = 0; m < nmem; ++m) {
            const void* z = (q < avail) ? memchr(p + q, 0, avail - q) : NULL;
            if (!z) return -1;
            const size_t len = (size_t)((const unsigned char*)z - (p + q));
            if (len > 31) return -1;
            memcpy(t->mname[m], p + q, len + 1);
            q += (len + 1 + 7) / 8 * 8;
            if (q + 32 > avail || p[q + 4] != 0) return -1;           /* byte offset, dimensionality 0 (no array members) */
            t->moff[m] = (size_t)le(p + q, 4);
            q += 32;
            h5t mt; size_t mu;
            if (rd_datatype(p + q, avail - q, &mt, &mu) || mt.kind == 6 || t->moff[m] + mt.size > t->size) return -1;
            t->mkind[m] = mt.kind;
            q += mu;
        }
        *used = q;
        return 0;
    }
    return -1;
}
static int rd_dataspace(const unsigned char* p, size_t avail, int* rank, hsize_t* dims) {
    if (avail < 8 || p[0] != 1 || p[2] != 0 || p[1] < 1 || p[1] > MAXRANK || avail < 8 + 8u * p[1]) return -1;
    *rank = p[1];
    for (int d = 0; d < *rank; ++d) dims[d] = le(p + 8 + 8 * d, 8);
    return 0;
}
static int rd_object(h5rd* r, uint64_t at, h5node* n);
static int rd_name(h5rd* r, uint64_t heap_data, uint64_t heap_size, uint64_t off, char* out, size_t cap) {
    if (off >= heap_size) return -1;
    size_t n = (size_t)(heap_size - off < cap ? heap_size - off : cap);
    if (rd_at(r, heap_data + off, out, n)) return -1;
    return memchr(out, 0, n) ? 0 : -1;
}
static int rd_snod(h5rd* r, uint64_t at, uint64_t heap_data, uint64_t heap_size, h5node* g) {
    unsigned char hd[8];
    if (rd_at(r, at, hd, 8) || memcmp(hd, "SNOD", 4) || hd[4] != 1) return -1;
    const unsigned cnt = (unsigned)le(hd + 6, 2);
    if (cnt > 2 * r->leaf_k) return -1;
    for (unsigned i = 0; i < cnt; ++i) {
        unsigned char e[40];
        char name[256];
        if (rd_at(r, at + 8 + 40ull * i, e, 40) || rd_name(r, heap_data, heap_size, le(e, 8), name, sizeof name)) return -1;
        if (find_child(g, name, strlen(name))) return -1;
        h5node* c = add_child(g, name, strlen(name), 0);
        if (rd_object(r, le(e + 8, 8), c)) return -1;
    }
    return 0;
}
static int rd_btree(h5rd* r, uint64_t at, uint64_t heap_data, uint64_t heap_size, h5node* g, int level_expect) {
    unsigned char hd[24];
    if (rd_at(r, at, hd, 24) || memcmp(hd, "TREE", 4) || hd[4] != 0) return -1;
    const int level = hd[5];
    const unsigned used = (unsigned)le(hd + 6, 2);
    if (used > 2 * r->int_k || level > 8 || (level_expect >= 0 && level != level_expect)) return -1;
    for (unsigned i = 0; i < used; ++i) {
        unsigned char kid[8];
        if (rd_at(r, at + 24 + 16ull * i + 8, kid, 8)) return -1;
        if (level > 0 ? rd_btree(r, le(kid, 8), heap_data, heap_size, g, level - 1) : rd_snod(r, le(kid, 8), heap_data, heap_size, g)) return -1;
    }
    return 0;
}
static int rd_object(h5rd* r, uint64_t at, h5node* n) {
    unsigned char hd[16];
    if (++r->depth > 16 || rd_at(r, at, hd, 16) || hd[0] != 1) return -1;
    const unsigned nmsg = (unsigned)le(hd + 2, 2);
    const uint64_t hsize = le(hd + 8, 4);
    if (hsize % 8 || hsize > (1u << 20)) return -1;
    unsigned char* b = (unsigned char*)malloc(hsize ? hsize : 1);
    int rc = rd_at(r, at + 16, b, hsize);
    int have_space = 0, have_type = 0, have_layout = 0;
    uint64_t stab_bt = UNDEF, stab_heap = UNDEF;
    size_t p = 0;
    unsigned seen = 0;
    while (!rc && p + 8 <= hsize) {
        const unsigned mtype = (unsigned)le(b + p, 2);
        const size_t msize = (size_t)le(b + p + 2, 2);
        const unsigned char* d = b + p + 8;
        if (msize % 8 || p + 8 + msize > hsize) { rc = -1; break; }
        if (mtype == 0x0001) { rc = rd_dataspace(d, msize, &n->rank, n->dims); have_space = 1; }
        else if (mtype == 0x0003) { size_t u; rc = rd_datatype(d, msize, &n->type, &u); have_type = 1; }
        else if (mtype == 0x0005) { if (msize < 4 || d[0] != 2) rc = -1; }
        else if (mtype == 0x0008) {
            if (msize < 18 || d[0] != 3 || d[1] != 1) rc = -1;                      /* version 3, contiguous */
            else { n->addr = le(d + 2, 8); n->bytes = le(d + 10, 8); have_layout = 1; }
        } else if (mtype == 0x0011) { if (msize < 16) rc = -1; else { stab_bt = le(d, 8); stab_heap = le(d + 8, 8); n->is_group = 1; } }
        else if (mtype == 0x000C) {                                                   /* attribute v1: kept so that the model is complete */
            if (msize < 8 || d[0] != 1) { rc = -1; break; }
            const size_t nsz = (size_t)le(d + 2, 2), tsz = (size_t)le(d + 4, 2), ssz = (size_t)le(d + 6, 2);
            size_t q = 8;
            h5attr a;
            memset(&a, 0, sizeof a);
            if (nsz < 1 || nsz > sizeof a.name || q + (nsz + 7) / 8 * 8 + (tsz + 7) / 8 * 8 + (ssz + 7) / 8 * 8 > msize || d[q + nsz - 1] != 0) { rc = -1; break; }
            memcpy(a.name, d + q, nsz);
            q += (nsz + 7) / 8 * 8;
            size_t u;
            if (rd_datatype(d + q, tsz, &a.type, &u) || u != tsz) { rc = -1; break; }
            q += (tsz + 7) / 8 * 8;
            if (rd_dataspace(d + q, ssz, &a.rank, a.dims)) { rc = -1; break; }
            q += (ssz + 7) / 8 * 8;
            a.bytes = a.type.size;
            for (int k = 0; k < a.rank; ++k) a.bytes *= (size_t)a.dims[k];
            if (q + a.bytes > msize) { rc = -1; break; }
            a.data = malloc(a.bytes ? a.bytes : 1);
            memcpy(a.data, d + q, a.bytes);
            n->attr = (h5attr*)realloc(n->attr, sizeof(h5attr) * (size_t)(n->nattr + 1));
            n->attr[n->nattr++] = a;
        } else if (mtype != 0) rc = -1;                                               /* nothing else is written by this library */
        p += 8 + msize;
        ++seen;
    }
    if (!rc && (p != hsize || seen != nmsg)) rc = -1;
    free(b);
    if (rc) return -1;
    if (n->is_group) {
        unsigned char hp[32];
        if (rd_at(r, stab_heap, hp, 32) || memcmp(hp, "HEAP", 4) || hp[4] != 0) return -1;
        if (rd_btree(r, stab_bt, le(hp + 24, 8), le(hp + 8, 8), n, -1)) return -1;
    } else {
        if (!have_space || !have_type || !have_layout) return -1;
        uint64_t want = n->type.size;
        for (int d = 0; d < n->rank; ++d) want *= n->dims[d];
        if (want != n->bytes || n->addr > r->eof || n->bytes > r->eof - n->addr) return -1;
    }
    --r->depth;
    return 0;
}
/* parses a file written by this library (or any HDF5 file restricted to the same structures) into the model */
static h5node* load_tree(FILE* fp) {
    unsigned char sb[96];
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (fseek(fp, 0, SEEK_SET) || fread(sb, 1, 96, fp) != 96 || memcmp(sb, sig, 8)) return NULL;
    if (sb[8] != 0 || sb[9] != 0 || sb[10] != 0 || sb[12] != 0 || sb[13] != 8 || sb[14] != 8) return NULL;   /* superblock v0, 8-byte offsets / lengths */
    h5rd r;
    memset(&r, 0, sizeof r);
    r.fp = fp;
    r.leaf_k = (unsigned)le(sb + 16, 2); r.int_k = (unsigned)le(sb + 18, 2);
    r.eof = le(sb + 40, 8);
    if (le(sb + 24, 8) != 0 || r.leaf_k == 0 || r.int_k == 0) return NULL;
    h5node* root = (h5node*)calloc(1, sizeof(h5node));
    root->name = strdup("");
    if (rd_object(&r, le(sb + 64, 8), root) || !root->is_group) { free_node(root); return NULL; }
    return root;
}

/* ------------------------------------------------------------------------------------------------ files */
static h5file* registry_find(const char* path) {
    for (int i = 0; i < 64; ++i) if (g_files[i] && !strcmp(g_files[i]->path, path)) return g_files[i];
    return NULL;
}
hid_t H5Fcreate(const char* name, unsigned flags, hid_t fcpl, hid_t fapl) {
    (void)flags; (void)fcpl; (void)fapl;
    h5file* f = registry_find(name);
    if (f) { if (f->open) return -1; free_node(f->root); }
    else {
        int slot = -1;
        for (int i = 0; i < 64; ++i) if (!g_files[i]) { slot = i; break; }
        if (slot < 0) return -1;
        f = (h5file*)calloc(1, sizeof *f);
        snprintf(f->path, sizeof f->path, "%s", name);
        g_files[slot] = f;
    }
    f->fp = fopen(name, "wb+");
    if (!f->fp) return -1;
    f->root = (h5node*)calloc(1, sizeof(h5node));
    f->root->name = strdup("");
    f->root->is_group = 1;
    f->data_end = 96;                      /* raw data start behind the superblock */
    f->open = 1;
    return new_id(K_FILE, f, f, f->root);
}
hid_t H5Fopen(const char* name, unsigned flags, hid_t fapl) {
    (void)fapl;
    h5file* f = registry_find(name);
    if (f) {                               /* a file this process created: reopened for the next save (hdf5_funcs.c:538) */
        if (f->open) return -1;
        f->fp = fopen(name, "rb+");
        if (!f->fp) return -1;
        f->open = 1;
        return new_id(K_FILE, f, f, f->root);
    }
    if (flags != H5F_ACC_RDONLY) return -1;        /* files of other processes are only read */
    FILE* fp = fopen(name, "rb");
    if (!fp) return -1;
    h5node* root = load_tree(fp);
    if (!root) { fclose(fp); return -1; }
    f = (h5file*)calloc(1, sizeof *f);
    snprintf(f->path, sizeof f->path, "%s", name);
    f->fp = fp; f->root = root; f->open = 1; f->readonly = 1;
    return new_id(K_FILE, f, f, f->root);
}
herr_t H5Fclose(hid_t file) {
    h5obj* o = get(file, K_FILE);
    if (!o) return -1;
    h5file* f = (h5file*)o->ptr;
    if (f->readonly) {                     /* not in the registry: the model goes with the handle */
        fclose(f->fp);
        free_node(f->root);
        free(f);
        drop(file);
        return 0;
    }
    int rc = flush_metadata(f);
    fclose(f->fp);
    f->fp = NULL;
    f->open = 0;
    drop(file);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ groups / links */
hid_t H5Gcreate(hid_t loc, const char* name, hid_t lcpl, hid_t gcpl, hid_t gapl) {
    (void)lcpl; (void)gcpl; (void)gapl;
    h5file* f; const char* last; size_t len;
    h5node* from = loc_node(loc, &f);
    if (!from) return -1;
    h5node* parent = walk(f, from, name, &last, &len);
    if (f->readonly || !parent || !len || find_child(parent, last, len)) return -1;
    return new_id(K_GROUP, add_child(parent, last, len, 1), f, NULL);
}
hid_t H5Gopen(hid_t loc, const char* name, hid_t gapl) {
    (void)gapl;
    h5file* f; const char* last; size_t len;
    h5node* from = loc_node(loc, &f);
    if (!from) return -1;
    h5node* parent = walk(f, from, name, &last, &len);
    h5node* g = parent ? (len ? find_child(parent, last, len) : parent) : NULL;
    if (!g || !g->is_group) return -1;
    return new_id(K_GROUP, g, f, NULL);
}
herr_t H5Gclose(hid_t group) { if (!get(group, K_GROUP)) return -1; drop(group); return 0; }
htri_t H5Lexists(hid_t loc, const char* name, hid_t lapl) {
    (void)lapl;
    h5file* f; const char* last; size_t len;
    h5node* from = loc_node(loc, &f);
    if (!from) return -1;
    h5node* parent = walk(f, from, name, &last, &len);
    return (parent && len && find_child(parent, last, len)) ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------------ dataspaces */
hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t* maxdims) {
    (void)maxdims;
    if (rank < 1 || rank > MAXRANK) return -1;
    h5space* s = (h5space*)calloc(1, sizeof *s);
    s->rank = rank;
    for (int d = 0; d < rank; ++d) { s->dims[d] = dims[d]; s->count[d] = dims[d]; }
    return new_id(K_SPACE, s, NULL, NULL);
}
herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t op, const hsize_t* start, const hsize_t* stride, const hsize_t* count, const hsize_t* block) {
    h5obj* o = get(space, K_SPACE);
    if (!o || op != H5S_SELECT_SET || stride || block) return -1;      /* start / count only */
    h5space* s = (h5space*)o->ptr;
    for (int d = 0; d < s->rank; ++d) {
        if (start[d] + count[d] > s->dims[d]) return -1;
        s->start[d] = start[d]; s->count[d] = count[d];
    }
    s->sel = 1;
    return 0;
}
herr_t H5Sclose(hid_t space) { h5obj* o = get(space, K_SPACE); if (!o) return -1; free(o->ptr); drop(space); return 0; }

/* ------------------------------------------------------------------------------------------------ datatypes */
hid_t H5Tcreate(H5T_class_t cls, size_t size) {
    if (cls != H5T_COMPOUND) return -1;
    h5t* t = (h5t*)calloc(1, sizeof *t);
    t->kind = 6; t->size = size;
    return new_id(K_TYPE, t, NULL, NULL);
}
herr_t H5Tinsert(hid_t type, const char* name, size_t offset, hid_t member) {
    h5obj* o = get(type, K_TYPE);
    h5t mt;
    if (!o || resolve_type(member, &mt) || mt.kind == 6) return -1;
    h5t* t = (h5t*)o->ptr;
    if (t->nmem == 8 || strlen(name) > 31 || offset + mt.size > t->size) return -1;
    snprintf(t->mname[t->nmem], 32, "%s", name);
    t->moff[t->nmem] = offset; t->mkind[t->nmem] = mt.kind;
    t->nmem++;
    return 0;
}
herr_t H5Tclose(hid_t type) { h5obj* o = get(type, K_TYPE); if (!o) return -1; free(o->ptr); drop(type); return 0; }

/* ------------------------------------------------------------------------------------------------ datasets */
hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl) {
    (void)lcpl; (void)dcpl; (void)dapl;
    h5file* f; const char* last; size_t len;
    h5node* from = loc_node(loc, &f);
    h5obj* so = get(space, K_SPACE);
    h5t t;
    if (!from || f->readonly || !so || resolve_type(type, &t)) return -1;
    h5node* parent = walk(f, from, name, &last, &len);
    if (!parent || !len || find_child(parent, last, len)) return -1;
    h5space* s = (h5space*)so->ptr;
    h5node* n = add_child(parent, last, len, 0);
    n->type = t; n->rank = s->rank;
    uint64_t bytes = t.size;
    for (int d = 0; d < s->rank; ++d) { n->dims[d] = s->dims[d]; bytes *= s->dims[d]; }
    n->bytes = bytes;
    n->addr = (f->data_end + 7) / 8 * 8;   /* contiguous storage, allocated now; unwritten parts read back as zeros */
    f->data_end = n->addr + bytes;
    return new_id(K_DSET, n, f, NULL);
}
/* elements that are contiguous in row-major order at the end of a box selection */
static uint64_t contiguous_run(const h5space* s) {
    uint64_t run = 1;
    for (int d = s->rank - 1; d >= 0; --d) {
        run *= s->count[d];
        if (s->count[d] != s->dims[d]) break;
    }
    return run;
}
static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
/* linear offset (elements) of the e-th selected element of a box selection */
static uint64_t sel_offset(const h5space* s, uint64_t e) {
    uint64_t off = 0, stride = 1;
    for (int d = s->rank - 1; d >= 0; --d) {
        const uint64_t c = e % s->count[d];
        e /= s->count[d];
        off += (s->start[d] + c) * stride;
        stride *= s->dims[d];
    }
    return off;
}
/* moves the selected elements between the dataset's contiguous storage and a buffer (no type conversion) */
static herr_t transfer(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, void* rbuf, const void* wbuf) {
    h5obj* o = get(dset, K_DSET);
    h5t mt;
    if (!o || (!rbuf && !wbuf) || resolve_type(mem_type, &mt)) return -1;
    h5node* n = (h5node*)o->ptr;
    h5file* f = o->file;
    if (wbuf && f->readonly) return -1;
    if (mt.size != n->type.size || mt.kind != n->type.kind) return -1;          /* no conversions */
    h5space whole;
    memset(&whole, 0, sizeof whole);
    whole.rank = n->rank;
    for (int d = 0; d < n->rank; ++d) { whole.dims[d] = n->dims[d]; whole.count[d] = n->dims[d]; }
    const h5space* fs = &whole;
    if (file_space != H5S_ALL) { h5obj* so = get(file_space, K_SPACE); if (!so) return -1; fs = (h5space*)so->ptr; }
    if (fs->rank != n->rank) return -1;
    for (int d = 0; d < n->rank; ++d) if (fs->dims[d] != n->dims[d]) return -1;
    uint64_t nsel = 1;
    for (int d = 0; d < fs->rank; ++d) nsel *= fs->count[d];
    h5space flat;
    const h5space* ms;
    if (mem_space != H5S_ALL) { h5obj* so = get(mem_space, K_SPACE); if (!so) return -1; ms = (h5space*)so->ptr; }
    else { memset(&flat, 0, sizeof flat); flat.rank = 1; flat.dims[0] = flat.count[0] = nsel; ms = &flat; }
    uint64_t msel = 1;
    for (int d = 0; d < ms->rank; ++d) msel *= ms->count[d];
    if (msel != nsel) return -1;
    if (nsel == 0) return 0;
    const uint64_t run = gcd64(contiguous_run(fs), contiguous_run(ms));
    const size_t es = n->type.size;
    for (uint64_t e = 0; e < nsel; e += run) {
        const uint64_t fo = sel_offset(fs, e), mo = sel_offset(ms, e);
        if (fseek(f->fp, (long)(n->addr + fo * es), SEEK_SET) != 0) return -1;
        if (wbuf) { if (fwrite((const char*)wbuf + mo * es, es, (size_t)run, f->fp) != (size_t)run) return -1; }
        else {
            /* storage that was allocated but never written lies beyond the data written so far: it reads back as zeros */
            const size_t got = fread((char*)rbuf + mo * es, es, (size_t)run, f->fp);
            if (got != (size_t)run) memset((char*)rbuf + (mo + got) * es, 0, ((size_t)run - got) * es);
        }
    }
    return 0;
}
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t dxpl, const void* buf) {
    (void)dxpl;
    return buf ? transfer(dset, mem_type, mem_space, file_space, NULL, buf) : -1;
}
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t dxpl, void* buf) {
    (void)dxpl;
    return buf ? transfer(dset, mem_type, mem_space, file_space, buf, NULL) : -1;
}
hid_t H5Dopen(hid_t loc, const char* name, hid_t dapl) {
    (void)dapl;
    h5file* f; const char* last; size_t len;
    h5node* from = loc_node(loc, &f);
    if (!from) return -1;
    h5node* parent = walk(f, from, name, &last, &len);
    h5node* n = (parent && len) ? find_child(parent, last, len) : NULL;
    if (!n || n->is_group) return -1;
    return new_id(K_DSET, n, f, NULL);
}
hid_t H5Dget_space(hid_t dset) {
    h5obj* o = get(dset, K_DSET);
    if (!o) return -1;
    const h5node* n = (const h5node*)o->ptr;
    return H5Screate_simple(n->rank, n->dims, NULL);
}
int H5Sget_simple_extent_ndims(hid_t space) { h5obj* o = get(space, K_SPACE); return o ? ((h5space*)o->ptr)->rank : -1; }
int H5Sget_simple_extent_dims(hid_t space, hsize_t* dims, hsize_t* maxdims) {
    h5obj* o = get(space, K_SPACE);
    if (!o) return -1;
    const h5space* s = (const h5space*)o->ptr;
    for (int d = 0; d < s->rank; ++d) { if (dims) dims[d] = s->dims[d]; if (maxdims) maxdims[d] = s->dims[d]; }
    return s->rank;
}
herr_t H5Dclose(hid_t dset) { if (!get(dset, K_DSET)) return -1; drop(dset); return 0; }

herr_t H5LTmake_dataset(hid_t loc, const char* name, int rank, const hsize_t* dims, hid_t type, const void* buffer) {
    hid_t sp = H5Screate_simple(rank, dims, NULL);
    if (sp < 0) return -1;
    hid_t ds = H5Dcreate(loc, name, type, sp, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
    herr_t rc = -1;
    if (ds >= 0) { rc = H5Dwrite(ds, type, H5S_ALL, H5S_ALL, H5P_DEFAULT, buffer); H5Dclose(ds); }
    H5Sclose(sp);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ attributes */
hid_t H5Acreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t acpl, hid_t aapl) {
    (void)acpl; (void)aapl;
    h5file* f = NULL;
    h5node* n = loc_node(loc, &f);
    if (!n) { h5obj* o = get(loc, K_DSET); if (o) { n = (h5node*)o->ptr; f = o->file; } }
    h5obj* so = get(space, K_SPACE);
    h5t t;
    if (!n || (f && f->readonly) || !so || resolve_type(type, &t) || strlen(name) > 63) return -1;
    for (int i = 0; i < n->nattr; ++i) if (!strcmp(n->attr[i].name, name)) return -1;
    n->attr = (h5attr*)realloc(n->attr, sizeof(h5attr) * (size_t)(n->nattr + 1));
    h5attr* a = &n->attr[n->nattr];
    memset(a, 0, sizeof *a);
    snprintf(a->name, sizeof a->name, "%s", name);
    a->type = t;
    const h5space* s = (h5space*)so->ptr;
    a->rank = s->rank;
    a->bytes = t.size;
    for (int d = 0; d < s->rank; ++d) { a->dims[d] = s->dims[d]; a->bytes *= (size_t)s->dims[d]; }
    a->data = calloc(1, a->bytes ? a->bytes : 1);
    return new_id(K_ATTR, n, f, (h5node*)(intptr_t)n->nattr++);
}
herr_t H5Awrite(hid_t attr, hid_t mem_type, const void* buf) {
    h5obj* o = get(attr, K_ATTR);
    h5t mt;
    if (!o || !buf || resolve_type(mem_type, &mt)) return -1;
    h5attr* a = &((h5node*)o->ptr)->attr[(int)(intptr_t)o->owner];
    if (mt.size != a->type.size || mt.kind != a->type.kind) return -1;
    memcpy(a->data, buf, a->bytes);
    return 0;
}
/* read side: attributes of a group, file root or dataset by name */
hid_t H5Aopen(hid_t loc, const char* name, hid_t aapl) {
    (void)aapl;
    h5file* f = NULL;
    h5node* n = loc_node(loc, &f);
    if (!n) { h5obj* o = get(loc, K_DSET); if (o) { n = (h5node*)o->ptr; f = o->file; } }
    if (!n) return -1;
    for (int i = 0; i < n->nattr; ++i)
        if (!strcmp(n->attr[i].name, name)) return new_id(K_ATTR, n, f, (h5node*)(intptr_t)i);
    return -1;
}
herr_t H5Aread(hid_t attr, hid_t mem_type, void* buf) {
    h5obj* o = get(attr, K_ATTR);
    h5t mt;
    if (!o || !buf || resolve_type(mem_type, &mt)) return -1;
    const h5attr* a = &((h5node*)o->ptr)->attr[(int)(intptr_t)o->owner];
    if (mt.size != a->type.size || mt.kind != a->type.kind) return -1;
    memcpy(buf, a->data, a->bytes);
    return 0;
}
herr_t H5Aclose(hid_t attr) { if (!get(attr, K_ATTR)) return -1; drop(attr); return 0; }

/* ------------------------------------------------------------------------------------------------ property lists */
hid_t H5Pcreate(hid_t cls) { (void)cls; return new_id(K_PLIST, NULL, NULL, NULL); }
herr_t H5Pclose(hid_t plist) { if (!get(plist, K_PLIST)) return -1; drop(plist); return 0; }
herr_t h5lite_noop(hid_t plist) { return get(plist, K_PLIST) ? 0 : -1; }
