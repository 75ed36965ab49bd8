/*
 * wasm_ref.c — a small WebAssembly (MVP + sign-ext + sat-trunc + bulk-memory) interpreter whose
 * only job is to run the reference's OWN compiled artefact, web/src/lib/pixo-wasm/pixo_bg.wasm
 * (pixo built by its authors with wasm-bindgen, `--features wasm,simd`), in the build container,
 * where no Rust toolchain and no wasm runtime exist.
 *
 * TEST INFRASTRUCTURE ONLY (lives under oracle/): it produces the golden fixtures under
 * tests/golden/ that pin the C oracle to real pixo output.  It is not part of the product and
 * is never used on the GPU box (the .wasm is read from /root/reference, which only exists in
 * the build container).  No reference code is copied: the .wasm is read where it lies.
 *
 * WebAssembly f32/f64 arithmetic is strict IEEE-754 (no fusion), so pixo-on-wasm computes the
 * same coefficients as pixo-on-x86.  Build: gcc -O2 -ffp-contract=off -fno-fast-math -msse2
 * -mfpmath=sse wasm_ref.c -lm -o wasm_ref
 *
 * usage:
 *   wasm_ref <pixo_bg.wasm> jpeg <in.raw> <w> <h> <color_type> <quality> <preset> <sub420> <out>
 *   wasm_ref <pixo_bg.wasm> png  <in.raw> <w> <h> <color_type> <preset> <lossy> <out>
 *   wasm_ref <pixo_bg.wasm> ops      (opcode census)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DIE(...) do { fprintf(stderr, "wasm_ref: " __VA_ARGS__); fprintf(stderr, "\n"); exit(2); } while (0)

typedef struct { uint32_t np, nr; uint8_t p[32], r[8]; } FuncType;
typedef struct {
    uint32_t type;
    const uint8_t *code, *end; /* body after locals */
    uint32_t nlocals;          /* declared locals (excluding params) */
    uint32_t code_off;         /* offset of body in module, for the control side tables */
} Func;

static uint8_t *mod; static size_t mod_len;
static FuncType *types; static uint32_t ntypes;
static Func *funcs; static uint32_t nfuncs, nimports;
static uint32_t *table; static uint32_t table_len;
static uint8_t *mem; static uint32_t mem_pages, mem_max_pages = 65536;
static uint64_t *globals; static uint32_t nglobals;
typedef struct { char name[64]; uint8_t kind; uint32_t idx; } Export;
static Export *exports; static uint32_t nexports;
typedef struct { const uint8_t *p; uint32_t len; int dropped; } DataSeg;
static DataSeg *datas; static uint32_t ndatas;
static uint32_t *ctl_end, *ctl_else; /* indexed by module offset of the block/loop/if opcode */
static uint64_t opcount[256], fc_count[32];
static char last_error[512];

static uint64_t leb_u(const uint8_t **pp)
{
    uint64_t r = 0; int s = 0; const uint8_t *p = *pp;
    for (;;) { uint8_t b = *p++; r |= (uint64_t)(b & 0x7f) << s; s += 7; if (!(b & 0x80)) break; }
    *pp = p; return r;
}
static int64_t leb_s(const uint8_t **pp)
{
    int64_t r = 0; int s = 0; const uint8_t *p = *pp; uint8_t b;
    do { b = *p++; r |= (int64_t)(b & 0x7f) << s; s += 7; } while (b & 0x80);
    if (s < 64 && (b & 0x40)) r |= -((int64_t)1 << s);
    *pp = p; return r;
}

/* skip the immediates of the instruction whose opcode byte was at p[-1]; returns new p */
static const uint8_t *skip_imm(uint8_t op, const uint8_t *p)
{
    switch (op) {
    case 0x02: case 0x03: case 0x04: leb_s(&p); break;               /* blocktype */
    case 0x0C: case 0x0D: leb_u(&p); break;
    case 0x0E: { uint64_t n = leb_u(&p); for (uint64_t i = 0; i <= n; i++) leb_u(&p); break; }
    case 0x10: leb_u(&p); break;
    case 0x11: leb_u(&p); leb_u(&p); break;
    case 0x1C: { uint64_t n = leb_u(&p); p += n; break; }
    case 0x20: case 0x21: case 0x22: case 0x23: case 0x24: case 0x25: case 0x26: leb_u(&p); break;
    case 0x3F: case 0x40: p++; break;
    case 0x41: leb_s(&p); break;
    case 0x42: leb_s(&p); break;
    case 0x43: p += 4; break;
    case 0x44: p += 8; break;
    case 0xD0: p++; break;
    case 0xD2: leb_u(&p); break;
    case 0xFC: {
        uint32_t sub = (uint32_t)leb_u(&p);
        if (sub < 32) fc_count[sub]++;
        switch (sub) {
        case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: break;
        case 8: leb_u(&p); p++; break;      /* memory.init seg, mem */
        case 9: leb_u(&p); break;           /* data.drop */
        case 10: p += 2; break;             /* memory.copy */
        case 11: p++; break;                /* memory.fill */
        case 12: leb_u(&p); leb_u(&p); break;
        case 13: leb_u(&p); break;
        case 14: leb_u(&p); leb_u(&p); break;
        case 15: case 16: case 17: leb_u(&p); break;
        default: DIE("unsupported 0xFC sub-opcode %u", sub);
        }
        break;
    }
    case 0xFD: DIE("SIMD (0xFD) opcodes are not supported by wasm_ref");
    default:
        if (op >= 0x28 && op <= 0x3E) { leb_u(&p); leb_u(&p); }  /* memarg */
        break;
    }
    return p;
}

static void prepass(Func *f)
{
    uint32_t stack[4096]; int sp = 0;
    const uint8_t *p = f->code;
    while (p < f->end) {
        uint32_t off = (uint32_t)(p - mod);
        uint8_t op = *p++;
        opcount[op]++;
        if (op == 0x02 || op == 0x03 || op == 0x04) {
            if (sp >= 4096) DIE("control nesting too deep");
            stack[sp++] = off; ctl_else[off] = 0;
        } else if (op == 0x05) {
            ctl_else[stack[sp - 1]] = off;
        } else if (op == 0x0B) {
            if (sp > 0) { ctl_end[stack[--sp]] = off; }
        }
        p = skip_imm(op, p);
    }
}

static void load_module(const char *path)
{
    FILE *fp = fopen(path, "rb"); if (!fp) DIE("cannot open %s", path);
    fseek(fp, 0, SEEK_END); mod_len = (size_t)ftell(fp); fseek(fp, 0, SEEK_SET);
    mod = malloc(mod_len); if (fread(mod, 1, mod_len, fp) != mod_len) DIE("read error"); fclose(fp);
    if (mod_len < 8 || memcmp(mod, "\0asm\1\0\0\0", 8)) DIE("not a wasm v1 module");
    ctl_end = calloc(mod_len, 4); ctl_else = calloc(mod_len, 4);
    const uint8_t *p = mod + 8, *endm = mod + mod_len;
    uint32_t *func_types = NULL; uint32_t ndefined = 0;
    while (p < endm) {
        uint8_t id = *p++; uint32_t sz = (uint32_t)leb_u(&p); const uint8_t *s = p, *se = p + sz;
        switch (id) {
        case 1: {
            ntypes = (uint32_t)leb_u(&s); types = calloc(ntypes, sizeof *types);
            for (uint32_t i = 0; i < ntypes; i++) {
                if (*s++ != 0x60) DIE("bad functype");
                types[i].np = (uint32_t)leb_u(&s); if (types[i].np > 32) DIE("too many params");
                for (uint32_t k = 0; k < types[i].np; k++) types[i].p[k] = *s++;
                types[i].nr = (uint32_t)leb_u(&s); if (types[i].nr > 8) DIE("too many results");
                for (uint32_t k = 0; k < types[i].nr; k++) types[i].r[k] = *s++;
            }
            break;
        }
        case 2: {
            uint32_t n = (uint32_t)leb_u(&s);
            funcs = calloc(n + 4096, sizeof *funcs);
            for (uint32_t i = 0; i < n; i++) {
                uint32_t l = (uint32_t)leb_u(&s); s += l; l = (uint32_t)leb_u(&s); s += l;
                uint8_t kind = *s++;
                if (kind != 0) DIE("only function imports are supported");
                funcs[nimports++].type = (uint32_t)leb_u(&s);
            }
            nfuncs = nimports;
            break;
        }
        case 3: {
            ndefined = (uint32_t)leb_u(&s); func_types = calloc(ndefined, 4);
            if (!funcs) funcs = calloc(ndefined + 16, sizeof *funcs);
            else funcs = realloc(funcs, (nimports + ndefined + 16) * sizeof *funcs);
            for (uint32_t i = 0; i < ndefined; i++) func_types[i] = (uint32_t)leb_u(&s);
            break;
        }
        case 4: {
            uint32_t n = (uint32_t)leb_u(&s); if (n != 1) DIE("expected one table");
            s++; uint8_t fl = *s++; table_len = (uint32_t)leb_u(&s); if (fl & 1) leb_u(&s);
            table = malloc(4 * (table_len + 1)); memset(table, 0xff, 4 * (table_len + 1));
            break;
        }
        case 5: {
            uint32_t n = (uint32_t)leb_u(&s); if (n != 1) DIE("expected one memory");
            uint8_t fl = *s++; mem_pages = (uint32_t)leb_u(&s); if (fl & 1) mem_max_pages = (uint32_t)leb_u(&s);
            mem = calloc((size_t)mem_pages, 65536);
            break;
        }
        case 6: {
            nglobals = (uint32_t)leb_u(&s); globals = calloc(nglobals, 8);
            for (uint32_t i = 0; i < nglobals; i++) {
                s += 2; uint8_t op = *s++;
                if (op == 0x41) globals[i] = (uint32_t)leb_s(&s);
                else if (op == 0x42) globals[i] = (uint64_t)leb_s(&s);
                else if (op == 0x43) { uint32_t v; memcpy(&v, s, 4); s += 4; globals[i] = v; }
                else if (op == 0x44) { memcpy(&globals[i], s, 8); s += 8; }
                else DIE("unsupported global initialiser");
                if (*s++ != 0x0B) DIE("bad global init");
            }
            break;
        }
        case 7: {
            nexports = (uint32_t)leb_u(&s); exports = calloc(nexports, sizeof *exports);
            for (uint32_t i = 0; i < nexports; i++) {
                uint32_t l = (uint32_t)leb_u(&s); memcpy(exports[i].name, s, l < 63 ? l : 63); s += l;
                exports[i].kind = *s++; exports[i].idx = (uint32_t)leb_u(&s);
            }
            break;
        }
        case 9: {
            uint32_t n = (uint32_t)leb_u(&s);
            for (uint32_t i = 0; i < n; i++) {
                uint32_t fl = (uint32_t)leb_u(&s);
                if (fl != 0) DIE("unsupported element segment kind %u", fl);
                if (*s++ != 0x41) DIE("bad elem offset"); uint32_t off = (uint32_t)leb_s(&s); s++;
                uint32_t cnt = (uint32_t)leb_u(&s);
                for (uint32_t k = 0; k < cnt; k++) {
                    uint32_t fi = (uint32_t)leb_u(&s);
                    if (off + k >= table_len) DIE("elem out of range"); table[off + k] = fi;
                }
            }
            break;
        }
        case 10: {
            uint32_t n = (uint32_t)leb_u(&s); if (n != ndefined) DIE("code/function count mismatch");
            for (uint32_t i = 0; i < n; i++) {
                uint32_t bsz = (uint32_t)leb_u(&s); const uint8_t *b = s, *be = s + bsz;
                Func *f = &funcs[nimports + i]; f->type = func_types[i];
                uint32_t ngroups = (uint32_t)leb_u(&b), nl = 0;
                for (uint32_t g = 0; g < ngroups; g++) { nl += (uint32_t)leb_u(&b); b++; }
                f->nlocals = nl; f->code = b; f->end = be; s = be;
            }
            nfuncs = nimports + n;
            break;
        }
        case 11: {
            ndatas = (uint32_t)leb_u(&s); datas = calloc(ndatas, sizeof *datas);
            for (uint32_t i = 0; i < ndatas; i++) {
                uint32_t fl = (uint32_t)leb_u(&s);
                if (fl == 0 || fl == 2) {
                    if (fl == 2) leb_u(&s);
                    if (*s++ != 0x41) DIE("bad data offset"); uint32_t off = (uint32_t)leb_s(&s); s++;
                    uint32_t l = (uint32_t)leb_u(&s);
                    if ((uint64_t)off + l > (uint64_t)mem_pages * 65536) DIE("data segment out of range");
                    memcpy(mem + off, s, l); datas[i].p = s; datas[i].len = l; s += l;
                } else { uint32_t l = (uint32_t)leb_u(&s); datas[i].p = s; datas[i].len = l; s += l; }
            }
            break;
        }
        default: break; /* custom, start (none), datacount */
        }
        p = se;
    }
    for (uint32_t i = nimports; i < nfuncs; i++) prepass(&funcs[i]);
}

/* ---- interpreter --------------------------------------------------------------------- */
#define STACK_SLOTS (1u << 22)
static uint64_t *vstack; static uint32_t vsp;
static uint32_t call_depth;

static inline void mem_check(uint64_t a, uint32_t n)
{
    if (a + n > (uint64_t)mem_pages * 65536) DIE("out-of-bounds memory access at %llu (+%u)", (unsigned long long)a, n);
}
static inline float u2f(uint64_t v) { float f; uint32_t u = (uint32_t)v; memcpy(&f, &u, 4); return f; }
static inline uint64_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline double u2d(uint64_t v) { double d; memcpy(&d, &v, 8); return d; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline uint32_t rotl32(uint32_t x, uint32_t n) { n &= 31; return n ? (x << n) | (x >> (32 - n)) : x; }
static inline uint32_t rotr32(uint32_t x, uint32_t n) { n &= 31; return n ? (x >> n) | (x << (32 - n)) : x; }
static inline uint64_t rotl64(uint64_t x, uint64_t n) { n &= 63; return n ? (x << n) | (x >> (64 - n)) : x; }
static inline uint64_t rotr64(uint64_t x, uint64_t n) { n &= 63; return n ? (x >> n) | (x << (64 - n)) : x; }
static float wasm_fminf(float a, float b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? a : b; return a < b ? a : b; }
static float wasm_fmaxf(float a, float b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? b : a; return a > b ? a : b; }
static double wasm_fmin(double a, double b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? a : b; return a < b ? a : b; }
static double wasm_fmax(double a, double b) { if (a != a || b != b) return NAN; if (a == 0 && b == 0) return signbit(a) ? b : a; return a > b ? a : b; }

static void host_call(uint32_t fi)
{
    /* the single import: wbg.__wbg_Error_*(ptr, len) -> externref (JsError construction) */
    (void)fi;
    uint32_t len = (uint32_t)vstack[--vsp], ptr = (uint32_t)vstack[--vsp];
    uint32_t n = len < sizeof last_error - 1 ? len : (uint32_t)sizeof last_error - 1;
    mem_check(ptr, n); memcpy(last_error, mem + ptr, n); last_error[n] = 0;
    vstack[vsp++] = 1;
}

typedef struct { const uint8_t *target; uint32_t height, arity; uint8_t is_loop; } Label;

static void blocktype(int64_t bt, uint32_t *np, uint32_t *nr)
{
    if (bt == -64) { *np = 0; *nr = 0; }
    else if (bt < 0) { *np = 0; *nr = 1; }
    else { *np = types[bt].np; *nr = types[bt].nr; }
}

static void exec(uint32_t fi)
{
    if (fi < nimports) { host_call(fi); return; }
    if (++call_depth > 20000) DIE("call stack exhausted");
    Func *f = &funcs[fi]; FuncType *ft = &types[f->type];
    uint32_t base = vsp - ft->np;              /* locals start (params already on the stack) */
    if (vsp + f->nlocals + 4096 >= STACK_SLOTS) DIE("value stack exhausted");
    for (uint32_t i = 0; i < f->nlocals; i++) vstack[vsp++] = 0;
    uint64_t *L = vstack + base;
    uint32_t opbase = vsp;
    Label labels[512]; int nl = 0;
    const uint8_t *p = f->code;
#define PUSH(v) (vstack[vsp++] = (uint64_t)(v))
#define POP() (vstack[--vsp])
#define TOP() (vstack[vsp - 1])
#define BR(depth) do { \
        int d_ = (int)(depth); \
        if (d_ >= nl) goto do_return; \
        Label *lb = &labels[nl - 1 - d_]; \
        uint32_t ar = lb->arity; \
        if (vsp - ar != lb->height) memmove(vstack + lb->height, vstack + vsp - ar, ar * 8); \
        vsp = lb->height + ar; p = lb->target; \
        nl -= d_; if (!lb->is_loop) nl--; \
    } while (0)
    for (;;) {
        uint32_t off = (uint32_t)(p - mod);
        uint8_t op = *p++;
        switch (op) {
        case 0x00: DIE("unreachable executed (wasm trap)%s%s", last_error[0] ? ": " : "", last_error);
        case 0x01: break;
        case 0x02: case 0x03: {
            int64_t bt = leb_s(&p); uint32_t np, nr; blocktype(bt, &np, &nr);
            if (nl >= 512) DIE("label stack overflow");
            Label *lb = &labels[nl++]; lb->height = vsp - np; lb->is_loop = op == 0x03;
            if (op == 0x03) { lb->target = p; lb->arity = np; }
            else { lb->target = mod + ctl_end[off] + 1; lb->arity = nr; }
            break;
        }
        case 0x04: {
            int64_t bt = leb_s(&p); uint32_t np, nr; blocktype(bt, &np, &nr);
            uint32_t c = (uint32_t)POP();
            Label *lb = &labels[nl++]; lb->height = vsp - np; lb->is_loop = 0;
            lb->target = mod + ctl_end[off] + 1; lb->arity = nr;
            if (!c) {
                if (ctl_else[off]) p = mod + ctl_else[off] + 1;
                else { p = mod + ctl_end[off] + 1; nl--; }
            }
            break;
        }
        case 0x05: { /* reached the end of the then-arm: leave the if */
            Label *lb = &labels[nl - 1]; p = lb->target; nl--; break;
        }
        case 0x0B:
            if (nl == 0) goto do_return;
            nl--; break;
        case 0x0C: { uint32_t d = (uint32_t)leb_u(&p); BR(d); break; }
        case 0x0D: { uint32_t d = (uint32_t)leb_u(&p); if ((uint32_t)POP()) BR(d); break; }
        case 0x0E: {
            uint32_t n = (uint32_t)leb_u(&p); uint32_t idx = (uint32_t)POP(); uint32_t d = 0;
            for (uint32_t i = 0; i <= n; i++) { uint32_t t = (uint32_t)leb_u(&p); if (i == idx || i == n) { d = t; if (i == idx) break; } }
            BR(d); break;
        }
        case 0x0F: goto do_return;
        case 0x10: { uint32_t callee = (uint32_t)leb_u(&p); exec(callee); break; }
        case 0x11: {
            uint32_t ti = (uint32_t)leb_u(&p); leb_u(&p);
            uint32_t idx = (uint32_t)POP();
            if (idx >= table_len || table[idx] == 0xffffffffu) DIE("call_indirect: null/out-of-range entry %u", idx);
            uint32_t callee = table[idx];
            FuncType *a = &types[funcs[callee].type], *b = &types[ti];
            if (a->np != b->np || a->nr != b->nr || memcmp(a->p, b->p, a->np) || memcmp(a->r, b->r, a->nr))
                DIE("call_indirect: signature mismatch");
            exec(callee); break;
        }
        case 0x1A: vsp--; break;
        case 0x1C: { uint32_t n = (uint32_t)leb_u(&p); p += n; } /* fallthrough */
        case 0x1B: { uint32_t c = (uint32_t)POP(); uint64_t b = POP(); if (!c) TOP() = b; break; }
        case 0x20: PUSH(L[leb_u(&p)]); break;
        case 0x21: L[leb_u(&p)] = POP(); break;
        case 0x22: L[leb_u(&p)] = TOP(); break;
        case 0x23: PUSH(globals[leb_u(&p)]); break;
        case 0x24: globals[leb_u(&p)] = POP(); break;
#define LOAD(T, W, conv) { leb_u(&p); uint64_t o = leb_u(&p); uint64_t a = (uint32_t)POP() + o; mem_check(a, W); T v; memcpy(&v, mem + a, W); PUSH(conv); break; }
        case 0x28: LOAD(uint32_t, 4, v)
        case 0x29: LOAD(uint64_t, 8, v)
        case 0x2A: LOAD(uint32_t, 4, v)
        case 0x2B: LOAD(uint64_t, 8, v)
        case 0x2C: LOAD(int8_t, 1, (uint32_t)(int32_t)v)
        case 0x2D: LOAD(uint8_t, 1, v)
        case 0x2E: LOAD(int16_t, 2, (uint32_t)(int32_t)v)
        case 0x2F: LOAD(uint16_t, 2, v)
        case 0x30: LOAD(int8_t, 1, (uint64_t)(int64_t)v)
        case 0x31: LOAD(uint8_t, 1, v)
        case 0x32: LOAD(int16_t, 2, (uint64_t)(int64_t)v)
        case 0x33: LOAD(uint16_t, 2, v)
        case 0x34: LOAD(int32_t, 4, (uint64_t)(int64_t)v)
        case 0x35: LOAD(uint32_t, 4, v)
#define STORE(W) { leb_u(&p); uint64_t o = leb_u(&p); uint64_t v = POP(); uint64_t a = (uint32_t)POP() + o; mem_check(a, W); memcpy(mem + a, &v, W); break; }
        case 0x36: STORE(4)
        case 0x37: STORE(8)
        case 0x38: STORE(4)
        case 0x39: STORE(8)
        case 0x3A: STORE(1)
        case 0x3B: STORE(2)
        case 0x3C: STORE(1)
        case 0x3D: STORE(2)
        case 0x3E: STORE(4)
        case 0x3F: p++; PUSH(mem_pages); break;
        case 0x40: {
            p++; uint32_t d = (uint32_t)POP();
            if ((uint64_t)mem_pages + d > mem_max_pages || (uint64_t)mem_pages + d > 49152) { PUSH(0xffffffffu); break; }
            uint8_t *nm = realloc(mem, ((size_t)mem_pages + d) * 65536);
            if (!nm) { PUSH(0xffffffffu); break; }
            memset(nm + (size_t)mem_pages * 65536, 0, (size_t)d * 65536);
            mem = nm; PUSH(mem_pages); mem_pages += d; break;
        }
        case 0x41: PUSH((uint32_t)leb_s(&p)); break;
        case 0x42: PUSH((uint64_t)leb_s(&p)); break;
        case 0x43: { uint32_t v; memcpy(&v, p, 4); p += 4; PUSH(v); break; }
        case 0x44: { uint64_t v; memcpy(&v, p, 8); p += 8; PUSH(v); break; }
#define I32 (uint32_t)
#define S32 (int32_t)(uint32_t)
#define UN32(expr) { uint32_t a = I32 POP(); PUSH((uint32_t)(expr)); break; }
#define BIN32(expr) { uint32_t b = I32 POP(); uint32_t a = I32 POP(); PUSH((uint32_t)(expr)); break; }
#define BIN64(expr) { uint64_t b = POP(); uint64_t a = POP(); PUSH((uint64_t)(expr)); break; }
#define CMP64(expr) { uint64_t b = POP(); uint64_t a = POP(); PUSH((uint32_t)(expr)); break; }
#define BINF(expr) { float b = u2f(POP()); float a = u2f(POP()); PUSH(f2u(expr)); break; }
#define CMPF(expr) { float b = u2f(POP()); float a = u2f(POP()); PUSH((uint32_t)(expr)); break; }
#define BIND(expr) { double b = u2d(POP()); double a = u2d(POP()); PUSH(d2u(expr)); break; }
#define CMPD(expr) { double b = u2d(POP()); double a = u2d(POP()); PUSH((uint32_t)(expr)); break; }
        case 0x45: UN32(a == 0)
        case 0x46: BIN32(a == b)
        case 0x47: BIN32(a != b)
        case 0x48: BIN32(S32 a < S32 b)
        case 0x49: BIN32(a < b)
        case 0x4A: BIN32(S32 a > S32 b)
        case 0x4B: BIN32(a > b)
        case 0x4C: BIN32(S32 a <= S32 b)
        case 0x4D: BIN32(a <= b)
        case 0x4E: BIN32(S32 a >= S32 b)
        case 0x4F: BIN32(a >= b)
        case 0x50: { uint64_t a = POP(); PUSH((uint32_t)(a == 0)); break; }
        case 0x51: CMP64(a == b)
        case 0x52: CMP64(a != b)
        case 0x53: CMP64((int64_t)a < (int64_t)b)
        case 0x54: CMP64(a < b)
        case 0x55: CMP64((int64_t)a > (int64_t)b)
        case 0x56: CMP64(a > b)
        case 0x57: CMP64((int64_t)a <= (int64_t)b)
        case 0x58: CMP64(a <= b)
        case 0x59: CMP64((int64_t)a >= (int64_t)b)
        case 0x5A: CMP64(a >= b)
        case 0x5B: CMPF(a == b)
        case 0x5C: CMPF(a != b)
        case 0x5D: CMPF(a < b)
        case 0x5E: CMPF(a > b)
        case 0x5F: CMPF(a <= b)
        case 0x60: CMPF(a >= b)
        case 0x61: CMPD(a == b)
        case 0x62: CMPD(a != b)
        case 0x63: CMPD(a < b)
        case 0x64: CMPD(a > b)
        case 0x65: CMPD(a <= b)
        case 0x66: CMPD(a >= b)
        case 0x67: UN32(a ? __builtin_clz(a) : 32)
        case 0x68: UN32(a ? __builtin_ctz(a) : 32)
        case 0x69: UN32(__builtin_popcount(a))
        case 0x6A: BIN32(a + b)
        case 0x6B: BIN32(a - b)
        case 0x6C: BIN32(a * b)
        case 0x6D: { int32_t b = S32 POP(), a = S32 POP(); if (!b) DIE("integer divide by zero"); if (a == INT32_MIN && b == -1) DIE("integer overflow"); PUSH((uint32_t)(a / b)); break; }
        case 0x6E: { uint32_t b = I32 POP(), a = I32 POP(); if (!b) DIE("integer divide by zero"); PUSH(a / b); break; }
        case 0x6F: { int32_t b = S32 POP(), a = S32 POP(); if (!b) DIE("integer divide by zero"); PUSH((uint32_t)((a == INT32_MIN && b == -1) ? 0 : a % b)); break; }
        case 0x70: { uint32_t b = I32 POP(), a = I32 POP(); if (!b) DIE("integer divide by zero"); PUSH(a % b); break; }
        case 0x71: BIN32(a & b)
        case 0x72: BIN32(a | b)
        case 0x73: BIN32(a ^ b)
        case 0x74: BIN32(a << (b & 31))
        case 0x75: BIN32((uint32_t)(S32 a >> (b & 31)))
        case 0x76: BIN32(a >> (b & 31))
        case 0x77: BIN32(rotl32(a, b))
        case 0x78: BIN32(rotr32(a, b))
        case 0x79: { uint64_t a = POP(); PUSH((uint64_t)(a ? __builtin_clzll(a) : 64)); break; }
        case 0x7A: { uint64_t a = POP(); PUSH((uint64_t)(a ? __builtin_ctzll(a) : 64)); break; }
        case 0x7B: { uint64_t a = POP(); PUSH((uint64_t)__builtin_popcountll(a)); break; }
        case 0x7C: BIN64(a + b)
        case 0x7D: BIN64(a - b)
        case 0x7E: BIN64(a * b)
        case 0x7F: { int64_t b = (int64_t)POP(), a = (int64_t)POP(); if (!b) DIE("integer divide by zero"); if (a == INT64_MIN && b == -1) DIE("integer overflow"); PUSH((uint64_t)(a / b)); break; }
        case 0x80: { uint64_t b = POP(), a = POP(); if (!b) DIE("integer divide by zero"); PUSH(a / b); break; }
        case 0x81: { int64_t b = (int64_t)POP(), a = (int64_t)POP(); if (!b) DIE("integer divide by zero"); PUSH((uint64_t)((a == INT64_MIN && b == -1) ? 0 : a % b)); break; }
        case 0x82: { uint64_t b = POP(), a = POP(); if (!b) DIE("integer divide by zero"); PUSH(a % b); break; }
        case 0x83: BIN64(a & b)
        case 0x84: BIN64(a | b)
        case 0x85: BIN64(a ^ b)
        case 0x86: BIN64(a << (b & 63))
        case 0x87: BIN64((uint64_t)((int64_t)a >> (b & 63)))
        case 0x88: BIN64(a >> (b & 63))
        case 0x89: BIN64(rotl64(a, b))
        case 0x8A: BIN64(rotr64(a, b))
        case 0x8B: { float a = u2f(POP()); PUSH(f2u(fabsf(a))); break; }
        case 0x8C: { uint32_t a = I32 POP(); PUSH(a ^ 0x80000000u); break; }
        case 0x8D: { float a = u2f(POP()); PUSH(f2u(ceilf(a))); break; }
        case 0x8E: { float a = u2f(POP()); PUSH(f2u(floorf(a))); break; }
        case 0x8F: { float a = u2f(POP()); PUSH(f2u(truncf(a))); break; }
        case 0x90: { float a = u2f(POP()); PUSH(f2u(nearbyintf(a))); break; }
        case 0x91: { float a = u2f(POP()); PUSH(f2u(sqrtf(a))); break; }
        case 0x92: BINF(a + b)
        case 0x93: BINF(a - b)
        case 0x94: BINF(a * b)
        case 0x95: BINF(a / b)
        case 0x96: BINF(wasm_fminf(a, b))
        case 0x97: BINF(wasm_fmaxf(a, b))
        case 0x98: { uint32_t b = I32 POP(), a = I32 POP(); PUSH((a & 0x7fffffffu) | (b & 0x80000000u)); break; }
        case 0x99: { double a = u2d(POP()); PUSH(d2u(fabs(a))); break; }
        case 0x9A: { uint64_t a = POP(); PUSH(a ^ 0x8000000000000000ull); break; }
        case 0x9B: { double a = u2d(POP()); PUSH(d2u(ceil(a))); break; }
        case 0x9C: { double a = u2d(POP()); PUSH(d2u(floor(a))); break; }
        case 0x9D: { double a = u2d(POP()); PUSH(d2u(trunc(a))); break; }
        case 0x9E: { double a = u2d(POP()); PUSH(d2u(nearbyint(a))); break; }
        case 0x9F: { double a = u2d(POP()); PUSH(d2u(sqrt(a))); break; }
        case 0xA0: BIND(a + b)
        case 0xA1: BIND(a - b)
        case 0xA2: BIND(a * b)
        case 0xA3: BIND(a / b)
        case 0xA4: BIND(wasm_fmin(a, b))
        case 0xA5: BIND(wasm_fmax(a, b))
        case 0xA6: { uint64_t b = POP(), a = POP(); PUSH((a & 0x7fffffffffffffffull) | (b & 0x8000000000000000ull)); break; }
        case 0xA7: { uint64_t a = POP(); PUSH((uint32_t)a); break; }
#define TRUNC(CONV, lo, hi, src) { if (src != src) DIE("invalid conversion to integer"); if (!(src > lo && src < hi)) DIE("integer overflow in trunc"); PUSH(CONV(src)); break; }
#define TO_S32(x) (uint32_t)(int32_t)(x)
#define TO_U32(x) (uint32_t)(x)
#define TO_S64(x) (uint64_t)(int64_t)(x)
#define TO_U64(x) (uint64_t)(x)
        case 0xA8: { float a = u2f(POP()); TRUNC(TO_S32, -2147483904.0f, 2147483648.0f, a) }
        case 0xA9: { float a = u2f(POP()); TRUNC(TO_U32, -1.0f, 4294967296.0f, a) }
        case 0xAA: { double a = u2d(POP()); TRUNC(TO_S32, -2147483649.0, 2147483648.0, a) }
        case 0xAB: { double a = u2d(POP()); TRUNC(TO_U32, -1.0, 4294967296.0, a) }
        case 0xAC: { uint32_t a = I32 POP(); PUSH((uint64_t)(int64_t)(int32_t)a); break; }
        case 0xAD: { uint32_t a = I32 POP(); PUSH((uint64_t)a); break; }
        case 0xAE: { float a = u2f(POP()); TRUNC(TO_S64, -9223373136366403584.0f, 9223372036854775808.0f, a) }
        case 0xAF: { float a = u2f(POP()); TRUNC(TO_U64, -1.0f, 18446744073709551616.0f, a) }
        case 0xB0: { double a = u2d(POP()); TRUNC(TO_S64, -9223372036854777856.0, 9223372036854775808.0, a) }
        case 0xB1: { double a = u2d(POP()); TRUNC(TO_U64, -1.0, 18446744073709551616.0, a) }
        case 0xB2: { int32_t a = S32 POP(); PUSH(f2u((float)a)); break; }
        case 0xB3: { uint32_t a = I32 POP(); PUSH(f2u((float)a)); break; }
        case 0xB4: { int64_t a = (int64_t)POP(); PUSH(f2u((float)a)); break; }
        case 0xB5: { uint64_t a = POP(); PUSH(f2u((float)a)); break; }
        case 0xB6: { double a = u2d(POP()); PUSH(f2u((float)a)); break; }
        case 0xB7: { int32_t a = S32 POP(); PUSH(d2u((double)a)); break; }
        case 0xB8: { uint32_t a = I32 POP(); PUSH(d2u((double)a)); break; }
        case 0xB9: { int64_t a = (int64_t)POP(); PUSH(d2u((double)a)); break; }
        case 0xBA: { uint64_t a = POP(); PUSH(d2u((double)a)); break; }
        case 0xBB: { float a = u2f(POP()); PUSH(d2u((double)a)); break; }
        case 0xBC: { uint32_t a = I32 POP(); PUSH(a); break; }          /* i32.reinterpret_f32 */
        case 0xBD: break;                                                /* i64.reinterpret_f64 */
        case 0xBE: { uint32_t a = I32 POP(); PUSH(a); break; }          /* f32.reinterpret_i32 */
        case 0xBF: break;
        case 0xC0: { uint32_t a = I32 POP(); PUSH((uint32_t)(int32_t)(int8_t)a); break; }
        case 0xC1: { uint32_t a = I32 POP(); PUSH((uint32_t)(int32_t)(int16_t)a); break; }
        case 0xC2: { uint64_t a = POP(); PUSH((uint64_t)(int64_t)(int8_t)a); break; }
        case 0xC3: { uint64_t a = POP(); PUSH((uint64_t)(int64_t)(int16_t)a); break; }
        case 0xC4: { uint64_t a = POP(); PUSH((uint64_t)(int64_t)(int32_t)a); break; }
        case 0xFC: {
            uint32_t sub = (uint32_t)leb_u(&p);
            switch (sub) {
            case 0: { float a = u2f(POP()); int32_t r; if (a != a) r = 0; else if (a <= -2147483648.0f) r = INT32_MIN; else if (a >= 2147483648.0f) r = INT32_MAX; else r = (int32_t)a; PUSH((uint32_t)r); break; }
            case 1: { float a = u2f(POP()); uint32_t r; if (a != a || a <= 0.0f) r = 0; else if (a >= 4294967296.0f) r = UINT32_MAX; else r = (uint32_t)a; PUSH(r); break; }
            case 2: { double a = u2d(POP()); int32_t r; if (a != a) r = 0; else if (a <= -2147483648.0) r = INT32_MIN; else if (a >= 2147483647.0) r = INT32_MAX; else r = (int32_t)a; PUSH((uint32_t)r); break; }
            case 3: { double a = u2d(POP()); uint32_t r; if (a != a || a <= 0.0) r = 0; else if (a >= 4294967295.0) r = UINT32_MAX; else r = (uint32_t)a; PUSH(r); break; }
            case 4: { float a = u2f(POP()); int64_t r; if (a != a) r = 0; else if (a <= -9223372036854775808.0f) r = INT64_MIN; else if (a >= 9223372036854775808.0f) r = INT64_MAX; else r = (int64_t)a; PUSH((uint64_t)r); break; }
            case 5: { float a = u2f(POP()); uint64_t r; if (a != a || a <= 0.0f) r = 0; else if (a >= 18446744073709551616.0f) r = UINT64_MAX; else r = (uint64_t)a; PUSH(r); break; }
            case 6: { double a = u2d(POP()); int64_t r; if (a != a) r = 0; else if (a <= -9223372036854775808.0) r = INT64_MIN; else if (a >= 9223372036854775808.0) r = INT64_MAX; else r = (int64_t)a; PUSH((uint64_t)r); break; }
            case 7: { double a = u2d(POP()); uint64_t r; if (a != a || a <= 0.0) r = 0; else if (a >= 18446744073709551616.0) r = UINT64_MAX; else r = (uint64_t)a; PUSH(r); break; }
            case 8: {
                uint32_t seg = (uint32_t)leb_u(&p); p++;
                uint32_t n = I32 POP(), s = I32 POP(), d = I32 POP();
                if (seg >= ndatas || (uint64_t)s + n > (datas[seg].dropped ? 0 : datas[seg].len)) DIE("memory.init out of range");
                mem_check(d, n); memcpy(mem + d, datas[seg].p + s, n); break;
            }
            case 9: { uint32_t seg = (uint32_t)leb_u(&p); if (seg < ndatas) datas[seg].dropped = 1; break; }
            case 10: { p += 2; uint32_t n = I32 POP(), s = I32 POP(), d = I32 POP(); mem_check(s, n); mem_check(d, n); memmove(mem + d, mem + s, n); break; }
            case 11: { p++; uint32_t n = I32 POP(), v = I32 POP(), d = I32 POP(); mem_check(d, n); memset(mem + d, (int)v, n); break; }
            default: DIE("unsupported 0xFC %u at runtime", sub);
            }
            break;
        }
        default: DIE("unsupported opcode 0x%02x at module offset %u", op, off);
        }
    }
do_return: {
        uint32_t nr = ft->nr;
        if (vsp - nr != base) memmove(vstack + base, vstack + vsp - nr, nr * 8);
        vsp = base + nr;
        (void)opbase;
        call_depth--;
    }
}

static uint32_t find_export(const char *name)
{
    for (uint32_t i = 0; i < nexports; i++)
        if (exports[i].kind == 0 && !strcmp(exports[i].name, name)) return exports[i].idx;
    DIE("export %s not found", name);
}

static uint32_t call_n(const char *name, int nargs, const uint32_t *args, int want)
{
    uint32_t fi = find_export(name);
    for (int i = 0; i < nargs; i++) vstack[vsp++] = args[i];
    exec(fi);
    return want ? (uint32_t)vstack[--vsp] : 0;
}

static uint8_t *read_file(const char *path, size_t *len)
{
    FILE *fp = fopen(path, "rb"); if (!fp) DIE("cannot open %s", path);
    fseek(fp, 0, SEEK_END); *len = (size_t)ftell(fp); fseek(fp, 0, SEEK_SET);
    uint8_t *b = malloc(*len + 1); if (fread(b, 1, *len, fp) != *len) DIE("read error"); fclose(fp); return b;
}

int main(int argc, char **argv)
{
    if (argc < 3) DIE("usage: wasm_ref <pixo_bg.wasm> jpeg|png|ops ...");
    load_module(argv[1]);
    if (!strcmp(argv[2], "ops")) {
        for (int i = 0; i < 256; i++) if (opcount[i]) printf("0x%02x %llu\n", i, (unsigned long long)opcount[i]);
        for (int i = 0; i < 32; i++) if (fc_count[i]) printf("0xfc.%d %llu\n", i, (unsigned long long)fc_count[i]);
        printf("funcs %u (imports %u) types %u table %u mem_pages %u globals %u datas %u\n", nfuncs, nimports, ntypes, table_len, mem_pages, nglobals, ndatas);
        return 0;
    }
    vstack = malloc(sizeof(uint64_t) * STACK_SLOTS);
    int is_jpeg = !strcmp(argv[2], "jpeg");
    if ((is_jpeg && argc != 11) || (!is_jpeg && argc != 10)) DIE("bad argument count");
    size_t len; uint8_t *in = read_file(argv[3], &len);
    uint32_t a1[1] = {(uint32_t)-16};
    uint32_t retptr = call_n("__wbindgen_add_to_stack_pointer", 1, a1, 1);
    uint32_t a2[2] = {(uint32_t)len, 1};
    uint32_t ptr0 = len ? call_n("__wbindgen_export", 2, a2, 1) : 1;
    mem_check(ptr0, (uint32_t)len); memcpy(mem + ptr0, in, len);
    const char *outpath;
    if (is_jpeg) {
        uint32_t a[9] = {retptr, ptr0, (uint32_t)len, (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]), (uint32_t)atoi(argv[6]),
                         (uint32_t)atoi(argv[7]), (uint32_t)atoi(argv[8]), (uint32_t)atoi(argv[9])};
        call_n("encodeJpeg", 9, a, 0); outpath = argv[10];
    } else {
        uint32_t a[8] = {retptr, ptr0, (uint32_t)len, (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]), (uint32_t)atoi(argv[6]),
                         (uint32_t)atoi(argv[7]), (uint32_t)atoi(argv[8])};
        call_n("encodePng", 8, a, 0); outpath = argv[9];
    }
    uint32_t r[4]; mem_check(retptr, 16); memcpy(r, mem + retptr, 16);
    if (r[3]) { fprintf(stderr, "pixo error: %s\n", last_error); return 3; }
    mem_check(r[0], r[1]);
    FILE *fo = fopen(outpath, "wb"); if (!fo) DIE("cannot write %s", outpath);
    fwrite(mem + r[0], 1, r[1], fo); fclose(fo);
    return 0;
}
