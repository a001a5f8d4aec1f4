/*
 * fasta_oracle.c -- CPU restatement of annalam/seqkit's per-read FASTQ batch path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for seqkit_b200: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, link, load or execute it.  The product path (libseqkit_b200.so, the `fasta`
 * host binary) never calls into it and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference is a Rust crate with no tests, fixtures or golden
 * vectors on this path and no Rust toolchain exists in this image, so the
 * reference binary cannot be run here.  This oracle is a line-by-line restatement of
 * the cited Rust source; it is cross-checked against a second, independently written
 * restatement (oracle/restatement.py) and against the known-answer tests derived from
 * the source in SURVEY.md section 8c (tests/golden/).
 *
 * Third-party semantics restated here (crate versions are semver ranges, no Cargo.lock):
 *   regex ^1.0   : Regex::find of " BC:[ACGTNacgtn+]+" = leftmost match, greedy class run.
 *   Rust std     : BufRead::read_line (UTF-8 validating, '\n' inclusive),
 *                  str::trim_end / str::trim (Unicode White_Space), release-mode wrapping
 *                  u8 subtraction, process::exit(-1) -> status 255, panic -> status 101.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <setjmp.h>

#define ORC_EXIT_OK 0
#define ORC_EXIT_ERROR 255 /* error! macro: exit(-1), common.rs:11-16 */
#define ORC_EXIT_PANIC 101 /* Rust panic in main thread */

/* ------------------------------------------------------------------ buffers */
typedef struct {
    uint8_t *p;
    size_t n, cap;
} obuf;

static void ob_reserve(obuf *b, size_t add) {
    if (b->n + add <= b->cap) return;
    size_t nc = b->cap ? b->cap : 256;
    while (nc < b->n + add) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    if (!b->p) abort();
    b->cap = nc;
}
static void ob_put(obuf *b, const void *s, size_t n) {
    if (!n) return;
    ob_reserve(b, n);
    memcpy(b->p + b->n, s, n);
    b->n += n;
}
static void ob_putc(obuf *b, uint8_t c) { ob_put(b, &c, 1); }
static void ob_puts(obuf *b, const char *s) { ob_put(b, s, strlen(s)); }
static void ob_printf(obuf *b, const char *fmt, ...) {
    char tmp[512];
    va_list ap;
    va_start(ap, fmt);
    int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    if (k < 0) abort();
    if ((size_t)k < sizeof tmp) {
        ob_put(b, tmp, (size_t)k);
        return;
    }
    char *big = (char *)malloc((size_t)k + 1);
    va_start(ap, fmt);
    vsnprintf(big, (size_t)k + 1, fmt, ap);
    va_end(ap);
    ob_put(b, big, (size_t)k);
    free(big);
}

/* ------------------------------------------------------------------ process model */
typedef struct {
    obuf out; /* stdout */
    obuf err; /* stderr */
    jmp_buf jb;
    int exit_code;
} proc;

/* error! macro, common.rs:11-16: "ERROR: " + message + "\n" on stderr, exit status 255 */
static void fatal(proc *P, const char *fmt, ...) {
    char tmp[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    ob_puts(&P->err, "ERROR: ");
    ob_puts(&P->err, tmp);
    ob_putc(&P->err, '\n');
    P->exit_code = ORC_EXIT_ERROR;
    longjmp(P->jb, 1);
}
/* error!("<prefix>{line}") with a line of any length */
static void fatal_line(proc *P, const char *prefix, const uint8_t *line, size_t n) {
    ob_puts(&P->err, "ERROR: ");
    ob_puts(&P->err, prefix);
    ob_put(&P->err, line, n);
    ob_putc(&P->err, '\n');
    P->exit_code = ORC_EXIT_ERROR;
    longjmp(P->jb, 1);
}
/* Rust panic (unwrap / slice OOB / assert): message text is rustc-version dependent, so
 * only the status (101) and "stderr non-empty" are part of the contract. */
static void rust_panic(proc *P, const char *what) {
    ob_puts(&P->err, "thread 'main' panicked: ");
    ob_puts(&P->err, what);
    ob_putc(&P->err, '\n');
    P->exit_code = ORC_EXIT_PANIC;
    longjmp(P->jb, 1);
}

/* ------------------------------------------------------------------ UTF-8 / White_Space */
/* str::from_utf8 acceptance (what BufRead::read_line enforces, common.rs:106-112). */
static int utf8_valid(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        if (c >= 0xC2 && c <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return 0;
            i += 2;
        } else if (c >= 0xE0 && c <= 0xEF) {
            if (i + 2 >= n) return 0;
            uint8_t c1 = s[i + 1], c2 = s[i + 2];
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return 0;
            if (c == 0xE0 && c1 < 0xA0) return 0;  /* overlong */
            if (c == 0xED && c1 >= 0xA0) return 0; /* surrogates */
            i += 3;
        } else if (c >= 0xF0 && c <= 0xF4) {
            if (i + 3 >= n) return 0;
            uint8_t c1 = s[i + 1], c2 = s[i + 2], c3 = s[i + 3];
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return 0;
            if (c == 0xF0 && c1 < 0x90) return 0;
            if (c == 0xF4 && c1 >= 0x90) return 0;
            i += 4;
        } else
            return 0;
    }
    return 1;
}

/* Unicode White_Space (char::is_whitespace). */
static int is_ws_cp(uint32_t cp) {
    if ((cp >= 0x09 && cp <= 0x0D) || cp == 0x20) return 1;
    if (cp < 0x80) return 0;
    return cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
           cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
/* Decode the char that ENDS at s[n-1]; valid UTF-8 assumed. Returns its byte length. */
static size_t last_char(const uint8_t *s, size_t n, uint32_t *cp) {
    size_t k = 1;
    while (k < 4 && k < n && (s[n - k] & 0xC0) == 0x80) k++;
    const uint8_t *c = s + n - k;
    if (k == 1) *cp = c[0];
    else if (k == 2) *cp = ((uint32_t)(c[0] & 0x1F) << 6) | (c[1] & 0x3F);
    else if (k == 3) *cp = ((uint32_t)(c[0] & 0x0F) << 12) | ((uint32_t)(c[1] & 0x3F) << 6) | (c[2] & 0x3F);
    else *cp = ((uint32_t)(c[0] & 0x07) << 18) | ((uint32_t)(c[1] & 0x3F) << 12) | ((uint32_t)(c[2] & 0x3F) << 6) | (c[3] & 0x3F);
    return k;
}
static size_t first_char(const uint8_t *s, size_t n, uint32_t *cp) {
    uint8_t c = s[0];
    size_t k = c < 0x80 ? 1 : c < 0xE0 ? 2 : c < 0xF0 ? 3 : 4;
    if (k > n) k = n;
    if (k == 1) *cp = c;
    else if (k == 2) *cp = ((uint32_t)(s[0] & 0x1F) << 6) | (s[1] & 0x3F);
    else if (k == 3) *cp = ((uint32_t)(s[0] & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F);
    else *cp = ((uint32_t)(s[0] & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) | ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
    return k;
}
/* str::trim_end: returns the new byte length. */
static size_t trim_end_len(const uint8_t *s, size_t n) {
    while (n > 0) {
        uint32_t cp;
        size_t k = last_char(s, n, &cp);
        if (!is_ws_cp(cp)) break;
        n -= k;
    }
    return n;
}
/* str::trim_start: returns the number of leading bytes removed. */
static size_t trim_start_off(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint32_t cp;
        size_t k = first_char(s + i, n - i, &cp);
        if (!is_ws_cp(cp)) break;
        i += k;
    }
    return i;
}

/* ------------------------------------------------------------------ FileReader */
/* common.rs:83-112. The byte source is a memory buffer (the CLI wrapper below slurps the
 * file / gunzip pipe first); read_line semantics are preserved: through the next '\n'
 * inclusive, last line may lack it, len==0 -> false, invalid UTF-8 -> fatal. */
typedef struct {
    const uint8_t *p;
    size_t n, pos;
} lrd;
typedef struct {
    const uint8_t *s;
    size_t n;
} line_t;

static int read_line(proc *P, lrd *r, line_t *l) {
    l->s = r->p + r->pos; /* line.clear() */
    l->n = 0;
    if (r->pos >= r->n) return 0;
    const uint8_t *nl = (const uint8_t *)memchr(r->p + r->pos, '\n', r->n - r->pos);
    size_t len = nl ? (size_t)(nl - (r->p + r->pos)) + 1 : r->n - r->pos;
    if (!utf8_valid(r->p + r->pos, len)) fatal(P, "I/O error while reading from file."); /* common.rs:110 */
    l->n = len;
    r->pos += len;
    return 1;
}
static int starts_with(const line_t *l, char c) { return l->n > 0 && l->s[0] == (uint8_t)c; }

/* ------------------------------------------------------------------ trim by quality */
/* fasta_trim_by_quality.rs:10-50 */
static void run_trim(proc *P, const uint8_t *in, size_t n, unsigned min_baseq) {
    lrd f = {in, n, 0};
    line_t line, seq, qual;
    const uint8_t baseq_offset = 33; /* :14 */
    while (read_line(P, &f, &line)) { /* :19 */
        if (!starts_with(&line, '@')) fatal(P, "Invalid FASTQ format encountered."); /* :20-22 */
        ob_put(&P->out, line.s, line.n); /* :23 */
        read_line(P, &f, &seq);          /* :24 */
        read_line(P, &f, &line);         /* :25 */
        read_line(P, &f, &qual);         /* :26 */

        int32_t total = -50; /* :28 */
        int32_t lowest_total = total;
        size_t k = trim_end_len(qual.s, qual.n); /* :31 */
        size_t lowest_k = k;
        while (k > 0) { /* :33-42 */
            k -= 1;
            total += (int32_t)(uint8_t)(qual.s[k] - baseq_offset) - (int32_t)min_baseq; /* :35 wrapping u8 */
            if (total > 0) break;
            if (total < lowest_total) {
                lowest_total = total;
                lowest_k = k;
            }
        }
        if (lowest_k == 0) { /* :44-45 */
            ob_puts(&P->out, "N\n+\n!\n");
        } else { /* :47  &seq[..lowest_k], &qual[..lowest_k] */
            if (lowest_k > seq.n) rust_panic(P, "byte index out of range of seq (fasta_trim_by_quality.rs:47)");
            if (lowest_k < seq.n && (seq.s[lowest_k] & 0xC0) == 0x80)
                rust_panic(P, "byte index is not a char boundary in seq (fasta_trim_by_quality.rs:47)");
            /* lowest_k <= trim_end(qual).len() and sits on a char boundary of qual by construction
             * unless it splits a multi-byte char, which the k-loop can do: */
            if (lowest_k < qual.n && (qual.s[lowest_k] & 0xC0) == 0x80)
                rust_panic(P, "byte index is not a char boundary in qual (fasta_trim_by_quality.rs:47)");
            ob_put(&P->out, seq.s, lowest_k);
            ob_puts(&P->out, "\n+\n");
            ob_put(&P->out, qual.s, lowest_k);
            ob_putc(&P->out, '\n');
        }
    }
}

/* ------------------------------------------------------------------ mask by quality */
/* fasta_mask_by_quality.rs:11-47 */
static void run_mask(proc *P, const uint8_t *in, size_t n, unsigned min_baseq) {
    lrd f = {in, n, 0};
    line_t line, seq, bq;
    obuf output = {0}; /* the per-record `output` String, printed only at :45 */
    while (read_line(P, &f, &line)) {                                                 /* :20 */
        if (!starts_with(&line, '@')) fatal(P, "Invalid FASTQ format encountered."); /* :21-23 */
        output.n = 0;                    /* :25 */
        ob_put(&output, line.s, line.n); /* :26 */
        read_line(P, &f, &seq);          /* :28 */
        read_line(P, &f, &line);         /* :29 */
        read_line(P, &f, &bq);           /* :30 */
        if (seq.n && seq.s[seq.n - 1] == '\n') seq.n--; /* :32 */
        if (bq.n && bq.s[bq.n - 1] == '\n') bq.n--;     /* :33 */
        if (seq.n != bq.n)                              /* :35-37 */
            fatal(P, "Read sequence and base qualities are of different length.");
        /* :40-43 zip over chars; `qual as u8` truncates the code point */
        size_t i = 0, j = 0;
        while (i < seq.n && j < bq.n) {
            uint32_t cb, cq;
            size_t kb = first_char(seq.s + i, seq.n - i, &cb);
            size_t kq = first_char(bq.s + j, bq.n - j, &cq);
            uint8_t q = (uint8_t)((uint8_t)cq - 33u);
            if (q < min_baseq) ob_putc(&output, 'N');
            else ob_put(&output, seq.s + i, kb);
            i += kb;
            j += kq;
        }
        ob_puts(&output, "\n+\n"); /* :44 */
        ob_put(&output, bq.s, bq.n);
        ob_putc(&output, '\n');
        ob_put(&P->out, output.p, output.n); /* :45 */
    }
    free(output.p);
}

/* ------------------------------------------------------------------ add barcode */
/* fasta_add_barcode.rs:11-45 */
static void run_add_barcode(proc *P, const uint8_t *fq, size_t n, const uint8_t *bc, size_t nb) {
    lrd fastq = {fq, n, 0}, bfile = {bc, nb, 0};
    line_t header, barcode = {(const uint8_t *)"", 0}, line;
    for (;;) {
        read_line(P, &bfile, &header); /* :20 (return value ignored; header cleared at EOF) */
        if (starts_with(&header, '@')) { /* :21-24 */
            read_line(P, &bfile, &barcode);
            read_line(P, &bfile, &line);
            read_line(P, &bfile, &line);
        } else if (starts_with(&header, '>')) { /* :25-27 */
            read_line(P, &bfile, &barcode);
        } /* else: previous `barcode` String is kept */
        if (!read_line(P, &fastq, &header)) break; /* :29-31 */
        ob_put(&P->out, header.s, trim_end_len(header.s, header.n)); /* :33 */
        ob_puts(&P->out, " BC:");
        ob_put(&P->out, barcode.s, trim_end_len(barcode.s, barcode.n));
        ob_putc(&P->out, '\n');
        if (starts_with(&header, '@')) { /* :35-38 */
            for (int i = 0; i < 3; i++) {
                read_line(P, &fastq, &line);
                ob_put(&P->out, line.s, line.n);
            }
        } else if (starts_with(&header, '>')) { /* :39-40 */
            read_line(P, &fastq, &line);
            ob_put(&P->out, line.s, line.n);
        } else { /* :41-43 */
            fatal_line(P, "Invalid FASTQ line:\n", header.s, header.n);
        }
    }
}

/* ------------------------------------------------------------------ demultiplex */
typedef struct {
    obuf name, barcode;
    obuf output[2]; /* decompressed content of {name}_1.fq.gz/{name}_2.fq.gz or {name}.fq.gz */
    uint64_t total_reads;
} sample_t;

typedef struct {
    proc P;
    sample_t *samples;
    size_t n_samples;
    int paired;
    int outputs_created; /* files are created at sheet-load time unless dry-run (:77-87) */
    uint64_t total_reads, identified_reads;
} demux_t;

static int is_bc_class(uint8_t c) {
    switch (c) {
    case 'A': case 'C': case 'G': case 'T': case 'N':
    case 'a': case 'c': case 'g': case 't': case 'n': case '+':
        return 1;
    }
    return 0;
}
/* Regex::find(" BC:[ACGTNacgtn+]+"), fasta_demultiplex.rs:38,140,221: leftmost match, greedy. */
static int bc_regex_find(const uint8_t *s, size_t n, size_t *start, size_t *end) {
    for (size_t i = 0; i + 5 <= n; i++) {
        if (s[i] == ' ' && s[i + 1] == 'B' && s[i + 2] == 'C' && s[i + 3] == ':' && is_bc_class(s[i + 4])) {
            size_t e = i + 5;
            while (e < n && is_bc_class(s[e])) e++;
            *start = i;
            *end = e;
            return 1;
        }
    }
    return 0;
}
/* fasta_demultiplex.rs:269-277 */
static size_t barcode_diff(const uint8_t *observed, const uint8_t *candidate, size_t len) {
    size_t mm = 0;
    for (size_t k = 0; k < len; k++) {
        if (candidate[k] == 'N' || candidate[k] == 'U') continue;
        if (observed[k] != candidate[k]) mm++;
    }
    return mm;
}

typedef struct {
    obuf key;
    uint64_t count;
} extra_t;

/* fasta_demultiplex.rs:30-265 */
static void run_demux(demux_t *D, const uint8_t *sheet, size_t ns, const uint8_t *r1, size_t n1, const uint8_t *r2,
                      size_t n2, int has_r2, const uint8_t *i1, size_t ni1, int has_i1, const uint8_t *i2, size_t ni2,
                      int has_i2, uint64_t dry_run) {
    proc *P = &D->P;
    lrd fastq[2] = {{r1, n1, 0}, {r2, n2, 0}};
    int paired_end = has_r2 ? 1 : 0; /* :43-46 */
    D->paired = paired_end;
    lrd index_fastq[2];
    int n_index = 0; /* :49-55 */
    if (has_i1) { index_fastq[n_index].p = i1; index_fastq[n_index].n = ni1; index_fastq[n_index].pos = 0; n_index++; }
    if (has_i2) { index_fastq[n_index].p = i2; index_fastq[n_index].n = ni2; index_fastq[n_index].pos = 0; n_index++; }

    ob_puts(&P->err, "Reading sample sheet...\n"); /* :58 */
    lrd sheet_rd = {sheet, ns, 0};
    line_t line;
    size_t barcode_len = 0;
    size_t cap = 0;
    while (read_line(P, &sheet_rd, &line)) { /* :63-95 */
        if (starts_with(&line, '#')) continue; /* :64 */
        size_t off = trim_start_off(line.s, line.n); /* line.trim() :65 */
        const uint8_t *t = line.s + off;
        size_t tn = trim_end_len(t, line.n - off);
        /* split('\t'): cols[0], cols[1] */
        const uint8_t *tab = (const uint8_t *)memchr(t, '\t', tn);
        if (!tab) continue; /* cols.len() < 2 :66 */
        size_t name_n = (size_t)(tab - t);
        const uint8_t *c1 = tab + 1;
        size_t rest = tn - name_n - 1;
        const uint8_t *tab2 = (const uint8_t *)memchr(c1, '\t', rest);
        size_t bc_n = tab2 ? (size_t)(tab2 - c1) : rest;
        if (bc_n == 0) fatal(P, "Sample %.*s has no barcode.", (int)name_n, t); /* :68 */
        if (barcode_len == 0) barcode_len = bc_n; /* :69-73 */
        else if (bc_n != barcode_len) fatal(P, "Barcodes in sample sheet must all be of same length.");
        if (D->n_samples == cap) {
            cap = cap ? cap * 2 : 16;
            D->samples = (sample_t *)realloc(D->samples, cap * sizeof(sample_t));
        }
        sample_t *s = &D->samples[D->n_samples++];
        memset(s, 0, sizeof *s);
        ob_put(&s->name, t, name_n);
        ob_put(&s->barcode, c1, bc_n);
        if (dry_run == 0) D->outputs_created = 1; /* :77-87 */
    }
    for (size_t s = 0; s < D->n_samples; s++) /* :98-104 */
        for (size_t k = s + 1; k < D->n_samples; k++)
            if (D->samples[s].name.n == D->samples[k].name.n &&
                memcmp(D->samples[s].name.p, D->samples[k].name.p, D->samples[s].name.n) == 0)
                fatal(P, "Sample %.*s is listed multiple times in sample sheet.", (int)D->samples[s].name.n,
                      D->samples[s].name.p);

    ob_printf(&P->err, "Starting demultiplexing in %s end mode...\n", paired_end ? "paired" : "single"); /* :106 */
    uint64_t total_reads = 0, identified_reads = 0;
    extra_t *extra = NULL;
    size_t n_extra = 0, cap_extra = 0;

    line_t header;
    obuf hdr = {0}, barcode = {0}, umi = {0}, l2 = {0};
    while (read_line(P, &fastq[0], &header)) { /* :117 */
        if (!starts_with(&header, '@')) {       /* :118-120 */
            fatal_line(P, "Invalid FASTQ header line:\n", header.s, header.n);
        }
        hdr.n = 0;
        ob_put(&hdr, header.s, header.n);
        barcode.n = 0; /* :123 */
        if (n_index > 0) { /* :126-136 */
            for (int q = 0; q < n_index; q++) {
                if (barcode.n) ob_putc(&barcode, '+');
                read_line(P, &index_fastq[q], &line);
                if (!starts_with(&line, '@')) rust_panic(P, "assertion failed: line.starts_with('@')");
                read_line(P, &index_fastq[q], &line);
                ob_put(&barcode, line.s, trim_end_len(line.s, line.n));
                read_line(P, &index_fastq[q], &line);
                if (!starts_with(&line, '+')) rust_panic(P, "assertion failed: line.starts_with('+')");
                read_line(P, &index_fastq[q], &line);
            }
        } else { /* :138-146 */
            size_t st, en;
            if (!bc_regex_find(hdr.p, hdr.n, &st, &en)) fatal(P, "No BC:xxxx field found.");
            ob_put(&barcode, hdr.p + st + 4, en - st - 4);
            memmove(hdr.p + st, hdr.p + en, hdr.n - en); /* header.drain(start..end) */
            hdr.n -= en - st;
        }
        if (barcode.n != barcode_len) /* :148-150 */
            fatal(P, "Sequenced barcode %.*s is of different length (%zu nt) than barcodes in the sample sheet (%zu nt).",
                  (int)barcode.n, barcode.p, barcode.n, barcode_len);

        size_t best = 0, equally_fine = 0, lowest_diff = SIZE_MAX; /* :154-166 */
        for (size_t s = 0; s < D->n_samples; s++) {
            size_t diff = barcode_diff(barcode.p, D->samples[s].barcode.p, barcode_len);
            if (diff < lowest_diff) {
                lowest_diff = diff;
                best = s;
                equally_fine = s;
            } else if (diff == lowest_diff)
                equally_fine = s;
        }
        total_reads++; /* :169 */
        D->total_reads = total_reads;
        int write_read_out = 0;
        if (lowest_diff <= 1) { /* :172 */
            if (best == equally_fine) {
                identified_reads++;
                D->identified_reads = identified_reads;
                D->samples[best].total_reads++;
                write_read_out = !(dry_run > 0);
            } else { /* :184-188 */
                sample_t *a = &D->samples[best], *b = &D->samples[equally_fine];
                ob_printf(&P->err,
                          "WARNING: Sequenced barcode %.*s was an equally good match (%zu mismatches) for samples %.*s "
                          "(%.*s) and %.*s (%.*s), and was therefore not assigned to any sample.\n",
                          (int)barcode.n, barcode.p, lowest_diff, (int)a->name.n, a->name.p, (int)a->barcode.n,
                          a->barcode.p, (int)b->name.n, b->name.p, (int)b->barcode.n, b->barcode.p);
            }
        } else if (dry_run > 0) { /* :190-194 */
            size_t e;
            for (e = 0; e < n_extra; e++)
                if (extra[e].key.n == barcode.n && memcmp(extra[e].key.p, barcode.p, barcode.n) == 0) break;
            if (e == n_extra) {
                if (n_extra == cap_extra) {
                    cap_extra = cap_extra ? cap_extra * 2 : 64;
                    extra = (extra_t *)realloc(extra, cap_extra * sizeof(extra_t));
                }
                memset(&extra[n_extra], 0, sizeof(extra_t));
                ob_put(&extra[n_extra].key, barcode.p, barcode.n);
                n_extra++;
            }
            extra[e].count++;
        }

        if (write_read_out) { /* :196-238 */
            sample_t *sm = &D->samples[best];
            umi.n = 0; /* :200-203  zip over chars of (sheet barcode, observed barcode) */
            {
                size_t i = 0, j = 0;
                while (i < sm->barcode.n && j < barcode.n) {
                    uint32_t ca, cb;
                    size_t ka = first_char(sm->barcode.p + i, sm->barcode.n - i, &ca);
                    size_t kb = first_char(barcode.p + j, barcode.n - j, &cb);
                    if (ca == 'U') ob_put(&umi, barcode.p + j, kb);
                    i += ka;
                    j += kb;
                }
            }
            ob_put(&sm->output[0], hdr.p, trim_end_len(hdr.p, hdr.n)); /* :206 */
            if (umi.n) { ob_puts(&sm->output[0], " UMI:"); ob_put(&sm->output[0], umi.p, umi.n); } /* :207 */
            ob_putc(&sm->output[0], '\n');
            for (int q = 0; q < 3; q++) { /* :209-212 */
                read_line(P, &fastq[0], &line);
                ob_put(&sm->output[0], line.s, line.n);
            }
            if (paired_end) { /* :215-238 */
                read_line(P, &fastq[1], &line);
                l2.n = 0;
                ob_put(&l2, line.s, line.n);
                if (n_index == 0) { /* :219-227 */
                    size_t st, en;
                    if (bc_regex_find(l2.p, l2.n, &st, &en) && en > 0) {
                        memmove(l2.p + st, l2.p + en, l2.n - en);
                        l2.n -= en - st;
                    }
                }
                ob_put(&sm->output[1], l2.p, trim_end_len(l2.p, l2.n)); /* :229 */
                if (umi.n) { ob_puts(&sm->output[1], " UMI:"); ob_put(&sm->output[1], umi.p, umi.n); }
                ob_putc(&sm->output[1], '\n');
                for (int q = 0; q < 3; q++) {
                    read_line(P, &fastq[1], &line);
                    ob_put(&sm->output[1], line.s, line.n);
                }
            }
        } else { /* :239-246 */
            for (int q = 0; q < 3; q++) read_line(P, &fastq[0], &line);
            if (paired_end)
                for (int q = 0; q < 4; q++) read_line(P, &fastq[1], &line);
        }
        D->total_reads = total_reads;
        D->identified_reads = identified_reads;
        if (dry_run > 0 && total_reads >= dry_run) break; /* :248 */
    }
    D->total_reads = total_reads;
    D->identified_reads = identified_reads;

    if (dry_run > 0) { /* :251-261 */
        ob_printf(&P->err, "Dry run completed with %llu clusters. Barcodes found:\n", (unsigned long long)total_reads);
        /* entries = samples (sheet order) ++ extra (HashMap order: arbitrary in the reference;
         * insertion order here); stable sort ascending by count, then reverse. */
        size_t ne = D->n_samples + n_extra;
        typedef struct { const uint8_t *p; size_t n; uint64_t c; } ent;
        ent *E = (ent *)malloc((ne ? ne : 1) * sizeof(ent));
        for (size_t s = 0; s < D->n_samples; s++) { E[s].p = D->samples[s].name.p; E[s].n = D->samples[s].name.n; E[s].c = D->samples[s].total_reads; }
        for (size_t e = 0; e < n_extra; e++) { E[D->n_samples + e].p = extra[e].key.p; E[D->n_samples + e].n = extra[e].key.n; E[D->n_samples + e].c = extra[e].count; }
        for (size_t a = 1; a < ne; a++) { /* stable insertion sort ascending */
            ent x = E[a];
            size_t b = a;
            while (b > 0 && E[b - 1].c > x.c) { E[b] = E[b - 1]; b--; }
            E[b] = x;
        }
        if (ne < 100) { /* &entries[0..100] :258 */
            free(E);
            rust_panic(P, "range end index 100 out of range for slice (fasta_demultiplex.rs:258)");
        }
        for (size_t q = 0; q < 100; q++) {
            ent *x = &E[ne - 1 - q];
            ob_puts(&P->out, "- ");
            ob_put(&P->out, x->p, x->n);
            ob_printf(&P->out, ": %llu\n", (unsigned long long)x->c);
        }
        free(E);
    }
    /* :263-264 */
    if (total_reads == 0)
        ob_printf(&P->err, "%llu / %llu (NaN%%) clusters carried a barcode matching one of the provided samples.\n",
                  (unsigned long long)identified_reads, (unsigned long long)total_reads);
    else
        ob_printf(&P->err, "%llu / %llu (%.1f%%) clusters carried a barcode matching one of the provided samples.\n",
                  (unsigned long long)identified_reads, (unsigned long long)total_reads,
                  (double)identified_reads / (double)total_reads * 100.0);
    free(hdr.p); free(barcode.p); free(umi.p); free(l2.p);
    for (size_t e = 0; e < n_extra; e++) free(extra[e].key.p);
    free(extra);
}


/* ================================================================== SURVEY.md section 8(f): the operators next to the path */
/* error! with a message of any length (fasta_check.rs quotes up to ten lines) */
static void fatal_buf(proc *P, const obuf *msg) {
    ob_puts(&P->err, "ERROR: ");
    ob_put(&P->err, msg->p, msg->n);
    ob_putc(&P->err, '\n');
    P->exit_code = ORC_EXIT_ERROR;
    longjmp(P->jb, 1);
}
static int char_boundary(const line_t *l, size_t k) { return k >= l->n || (l->s[k] & 0xC0) != 0x80; }

/* ------------------------------------------------------------------ trim --first / --last */
/* fasta_trim.rs:14-48 */
static void run_trim_fixed(proc *P, const uint8_t *in, size_t n, size_t remove_first, size_t remove_last) {
    lrd f = {in, n, 0};
    line_t line, seq, qual;
    while (read_line(P, &f, &line)) { /* :27 */
        if (!starts_with(&line, '>') && !starts_with(&line, '@')) /* :28-30 */
            fatal(P, "Invalid FASTA/FASTQ format encountered.");
        read_line(P, &f, &seq);                          /* :32 */
        size_t seq_len = trim_end_len(seq.s, seq.n);     /* :33 */
        int cut = remove_first + remove_last < seq_len;  /* :34 */
        if (cut) { /* :35 print!("{}{}\n", line, &seq[remove_first..seq_len-remove_last]) */
            if (!char_boundary(&seq, remove_first) || !char_boundary(&seq, seq_len - remove_last))
                rust_panic(P, "byte index is not a char boundary in seq (fasta_trim.rs:35)");
            ob_put(&P->out, line.s, line.n);
            ob_put(&P->out, seq.s + remove_first, seq_len - remove_last - remove_first);
            ob_putc(&P->out, '\n');
        } else { /* :37 */
            ob_put(&P->out, line.s, line.n);
            ob_putc(&P->out, '\n');
        }
        if (starts_with(&line, '@')) { /* :40 */
            read_line(P, &f, &line);   /* :41 */
            read_line(P, &f, &qual);   /* :42 */
            if (cut) { /* :44 print!("+\n{}\n", &qual[remove_first..seq_len-remove_last]): the slice is evaluated first */
                if (seq_len - remove_last > qual.n) rust_panic(P, "byte index out of range of qual (fasta_trim.rs:44)");
                if (!char_boundary(&qual, remove_first) || !char_boundary(&qual, seq_len - remove_last))
                    rust_panic(P, "byte index is not a char boundary in qual (fasta_trim.rs:44)");
                ob_puts(&P->out, "+\n");
                ob_put(&P->out, qual.s + remove_first, seq_len - remove_last - remove_first);
                ob_putc(&P->out, '\n');
            } else { /* :46 */
                ob_puts(&P->out, "+\n\n");
            }
        }
    }
}

/* ------------------------------------------------------------------ check */
/* fasta_check.rs:14-47: a reader that remembers the last ten lines */
typedef struct {
    lrd file;
    size_t lines_read;
    line_t prev[10];
    int n_prev;
} memrd;
static int mem_read_line(proc *P, memrd *r, line_t *l) {
    if (!read_line(P, &r->file, l)) return 0; /* :31 (the line is cleared) */
    if (r->n_prev == 10) {                    /* :32-35 */
        memmove(r->prev, r->prev + 1, 9 * sizeof(line_t));
        r->n_prev = 9;
    }
    r->prev[r->n_prev++] = *l;
    r->lines_read++; /* :36 */
    return 1;
}
static void check_fail(proc *P, memrd *r, const char *what) { /* :40-46 history(), :59-60 / :64-65 */
    obuf m = {0};
    ob_printf(&m, "%s on line %zu:\n", what, r->lines_read);
    for (int k = 0; k < r->n_prev; k++) {
        ob_put(&m, r->prev[k].s, r->prev[k].n);
        ob_putc(&m, '\n');
    }
    ob_putc(&m, '\n');
    fatal_buf(P, &m);
}
/* fasta_check.rs:49-70 */
static void run_check(proc *P, const uint8_t *in, size_t n) {
    memrd r;
    memset(&r, 0, sizeof r);
    r.file.p = in;
    r.file.n = n;
    line_t line;
    while (mem_read_line(P, &r, &line)) {  /* :54 */
        if (starts_with(&line, '>')) {     /* :55-56 */
            mem_read_line(P, &r, &line);
        } else if (starts_with(&line, '@')) { /* :57-63 */
            mem_read_line(P, &r, &line);
            mem_read_line(P, &r, &line);
            if (!starts_with(&line, '+')) check_fail(P, &r, "Missing quality header prefix '+'");
            mem_read_line(P, &r, &line);
        } else { /* :64-66 */
            check_fail(P, &r, "Missing header prefix '>' or '@'");
        }
    }
}

/* ------------------------------------------------------------------ statistics */
static int stat_class(uint8_t c) { /* fasta_statistics.rs:17: [ACGTNacgtn], no '+' */
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': return 1;
    }
    return 0;
}
typedef struct {
    uint8_t *s;
    size_t n;
    uint64_t count;
} stat_ent;
static int stat_cmp_key(const void *a, const void *b) {
    const stat_ent *x = (const stat_ent *)a, *y = (const stat_ent *)b;
    size_t m = x->n < y->n ? x->n : y->n;
    int c = memcmp(x->s, y->s, m);
    if (c) return c;
    return x->n < y->n ? -1 : x->n > y->n;
}
static int stat_cmp_rank(const void *a, const void *b) { /* count descending; equal counts: barcode descending */
    const stat_ent *x = (const stat_ent *)a, *y = (const stat_ent *)b;
    if (x->count != y->count) return x->count > y->count ? -1 : 1;
    return -stat_cmp_key(a, b);
}
/* fasta_statistics.rs:13-52.  The reference collects a HashMap, sorts by count (stable) and reverses: entries of
 * equal count come out in an order that depends on the hash seed.  The restatement fixes that order (barcode
 * descending), and so does the product; everything else is the reference's. */
static void run_statistics(proc *P, const uint8_t *in, size_t n) {
    lrd f = {in, n, 0};
    line_t line;
    uint64_t total_records = 0;
    stat_ent *ents = NULL;
    size_t n_ents = 0, cap = 0;
    while (read_line(P, &f, &line)) { /* :24 */
        /* :26-29 leftmost match of " BC:[ACGTNacgtn]+" */
        for (size_t k = 0; k + 5 <= line.n; k++)
            if (line.s[k] == ' ' && line.s[k + 1] == 'B' && line.s[k + 2] == 'C' && line.s[k + 3] == ':' && stat_class(line.s[k + 4])) {
                size_t e = k + 5;
                while (e < line.n && stat_class(line.s[e])) e++;
                if (n_ents == cap) {
                    cap = cap ? cap * 2 : 1024;
                    ents = (stat_ent *)realloc(ents, cap * sizeof(stat_ent));
                }
                ents[n_ents].s = (uint8_t *)line.s + k + 4;
                ents[n_ents].n = e - k - 4;
                ents[n_ents].count = 1;
                n_ents++;
                break;
            }
        if (starts_with(&line, '@')) { /* :32-33 */
            line_t t;
            for (int i = 0; i < 3; i++) read_line(P, &f, &t);
        } else if (starts_with(&line, '>')) { /* :34-35 */
            line_t t;
            read_line(P, &f, &t);
        } else { /* :36-38 */
            obuf m = {0};
            ob_puts(&m, "Invalid FASTQ header:\n");
            ob_put(&m, line.s, line.n);
            fatal_buf(P, &m);
        }
        total_records += 1; /* :40 */
    }
    ob_printf(&P->out, "Total sequence records: %llu\n", (unsigned long long)total_records); /* :43 */
    ob_puts(&P->out, "Most frequent sample barcodes:\n");                                   /* :45 */
    /* the HashMap: merge equal barcodes */
    qsort(ents, n_ents, sizeof(stat_ent), stat_cmp_key);
    size_t m = 0;
    for (size_t i = 0; i < n_ents; i++) {
        if (m && stat_cmp_key(&ents[m - 1], &ents[i]) == 0) ents[m - 1].count++;
        else ents[m++] = ents[i];
    }
    qsort(ents, m, sizeof(stat_ent), stat_cmp_rank); /* :46-49 */
    if (m < 100) rust_panic(P, "range end index 100 out of range for slice (fasta_statistics.rs:50)");
    for (size_t i = 0; i < 100; i++) { /* :50-52 */
        ob_puts(&P->out, "- ");
        ob_put(&P->out, ents[i].s, ents[i].n);
        ob_printf(&P->out, ": %llu\n", (unsigned long long)ents[i].count);
    }
    free(ents);
}

/* ------------------------------------------------------------------ interleave / deinterleave */
static void not_fastx(proc *P, const line_t *line) { /* "Line is not FASTA/FASTQ format: {}" */
    obuf m = {0};
    ob_puts(&m, "Line is not FASTA/FASTQ format: ");
    ob_put(&m, line->s, line->n);
    fatal_buf(P, &m);
}
/* fasta_interleave.rs:14-35 */
static void run_interleave(proc *P, const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    lrd f1 = {a, na, 0}, f2 = {b, nb, 0};
    line_t line;
    while (read_line(P, &f1, &line)) { /* :20 */
        int lines = starts_with(&line, '@') ? 4 : starts_with(&line, '>') ? 2 : 0; /* :21-23 */
        if (!lines) not_fastx(P, &line);
        ob_put(&P->out, line.s, line.n); /* :24 */
        for (int k = 0; k < lines - 1; k++) { /* :25-27 */
            read_line(P, &f1, &line);
            ob_put(&P->out, line.s, line.n);
        }
        read_line(P, &f2, &line); /* :29 */
        if ((lines == 4 && !starts_with(&line, '@')) || (lines == 2 && !starts_with(&line, '>'))) /* :30-33 */
            fatal(P, "Input files do not share a consistent format.");
        ob_put(&P->out, line.s, line.n); /* :34 */
        for (int k = 0; k < lines - 1; k++) {
            read_line(P, &f2, &line);
            ob_put(&P->out, line.s, line.n);
        }
    }
}
/* fasta_deinterleave.rs:14-39; the two gzip outputs are out / out2 */
static void run_deinterleave(proc *P, obuf *out2, const uint8_t *in, size_t n) {
    lrd f = {in, n, 0};
    line_t line;
    while (read_line(P, &f, &line)) { /* :22 */
        int lines = starts_with(&line, '@') ? 4 : starts_with(&line, '>') ? 2 : 0; /* :23-25 */
        if (!lines) not_fastx(P, &line);
        ob_put(&P->out, line.s, line.n); /* :26 */
        for (int k = 0; k < lines - 1; k++) {
            read_line(P, &f, &line);
            ob_put(&P->out, line.s, line.n);
        }
        read_line(P, &f, &line); /* :31 */
        if ((lines == 4 && !starts_with(&line, '@')) || (lines == 2 && !starts_with(&line, '>'))) /* :32-35 */
            fatal(P, "Interleaved FASTA records are not in consistent format.");
        ob_put(out2, line.s, line.n); /* :36 */
        for (int k = 0; k < lines - 1; k++) {
            read_line(P, &f, &line);
            ob_put(out2, line.s, line.n);
        }
    }
}

/* ------------------------------------------------------------------ extract dual umi */
/* fasta_extract_dual_umi.rs:14-72 */
static void run_dual_umi(proc *P, const uint8_t *in, size_t n, size_t first_bases) {
    lrd f = {in, n, 0};
    line_t header_1, header_2 = {(const uint8_t *)"", 0}, seq_1 = header_2, seq_2 = header_2, qual_1 = header_2, qual_2 = header_2, line;
    while (read_line(P, &f, &header_1)) { /* :30 */
        int fastq_format;
        if (starts_with(&header_1, '@')) fastq_format = 1; /* :33-35 */
        else if (starts_with(&header_1, '>')) fastq_format = 0;
        else {
            obuf m = {0};
            ob_puts(&m, "Header is not valid FASTA/FASTQ:\n");
            ob_put(&m, header_1.s, header_1.n);
            fatal_buf(P, &m);
            return;
        }
        if (fastq_format) { /* :37-47 */
            read_line(P, &f, &seq_1);
            read_line(P, &f, &line);
            read_line(P, &f, &qual_1);
            read_line(P, &f, &header_2);
            read_line(P, &f, &seq_2);
            read_line(P, &f, &line);
            read_line(P, &f, &qual_2);
            if (!starts_with(&header_2, '@')) fatal(P, "Invalid FASTQ record found in input file.");
        } else { /* :48-54 */
            read_line(P, &f, &seq_1);
            read_line(P, &f, &header_2);
            read_line(P, &f, &seq_2);
            if (!starts_with(&header_2, '>')) fatal(P, "Invalid FASTA record found in input file.");
        }
        /* :56-58 umi = seq_1[0..N] + "+" + seq_2[0..N] */
        if (first_bases > seq_1.n || !char_boundary(&seq_1, first_bases)) rust_panic(P, "byte index out of range of seq_1 (fasta_extract_dual_umi.rs:56)");
        if (first_bases > seq_2.n || !char_boundary(&seq_2, first_bases)) rust_panic(P, "byte index out of range of seq_2 (fasta_extract_dual_umi.rs:58)");
        if (fastq_format && (first_bases > qual_1.n || first_bases > qual_2.n || !char_boundary(&qual_1, first_bases) || !char_boundary(&qual_2, first_bases)))
            rust_panic(P, "byte index out of range of qual (fasta_extract_dual_umi.rs:63-65)");
        const line_t *hs[2] = {&header_1, &header_2}, *ss[2] = {&seq_1, &seq_2}, *qs[2] = {&qual_1, &qual_2};
        for (int m = 0; m < 2; m++) { /* :60-70, one print! per pair: all slices are taken before anything is written */
            ob_put(&P->out, hs[m]->s, trim_end_len(hs[m]->s, hs[m]->n));
            ob_puts(&P->out, " RX:");
            ob_put(&P->out, seq_1.s, first_bases);
            ob_putc(&P->out, '+');
            ob_put(&P->out, seq_2.s, first_bases);
            ob_putc(&P->out, '\n');
            ob_put(&P->out, ss[m]->s + first_bases, ss[m]->n - first_bases);
            if (fastq_format) {
                ob_puts(&P->out, "+\n");
                ob_put(&P->out, qs[m]->s + first_bases, qs[m]->n - first_bases);
            }
        }
    }
}

/* ------------------------------------------------------------------ exported C API (ctypes) */
typedef struct {
    int exit_code;
    uint8_t *out; size_t out_n;
    uint8_t *err; size_t err_n;
} orc_result;

static void finish(proc *P, orc_result *R) {
    R->exit_code = P->exit_code;
    R->out = P->out.p; R->out_n = P->out.n;
    R->err = P->err.p; R->err_n = P->err.n;
}

void orc_result_free(orc_result *R) {
    free(R->out); free(R->err);
    memset(R, 0, sizeof *R);
}

int orc_trim_by_quality(const uint8_t *in, size_t n, unsigned min_baseq, orc_result *R) {
    proc *P = (proc *)calloc(1, sizeof(proc));
    if (setjmp(P->jb) == 0) run_trim(P, in, n, min_baseq);
    finish(P, R);
    free(P);
    return R->exit_code;
}
int orc_mask_by_quality(const uint8_t *in, size_t n, unsigned min_baseq, orc_result *R) {
    proc *P = (proc *)calloc(1, sizeof(proc));
    if (setjmp(P->jb) == 0) run_mask(P, in, n, min_baseq);
    finish(P, R);
    free(P);
    return R->exit_code;
}
int orc_add_barcode(const uint8_t *fq, size_t n, const uint8_t *bc, size_t nb, orc_result *R) {
    proc *P = (proc *)calloc(1, sizeof(proc));
    if (setjmp(P->jb) == 0) run_add_barcode(P, fq, n, bc, nb);
    finish(P, R);
    free(P);
    return R->exit_code;
}

/* SURVEY.md section 8(f) operators: op 0 trim --first=x --last=y, 1 check, 2 statistics, 3 interleave (a, b), 4 deinterleave
 * (second output in out2), 5 extract dual umi --first-bases=x */
int orc_next(int op, const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint64_t x, uint64_t y, orc_result *R,
             uint8_t **out2, size_t *out2_n) {
    proc *P = (proc *)calloc(1, sizeof(proc));
    obuf *o2 = (obuf *)calloc(1, sizeof(obuf)); /* (on the heap: it is written between setjmp and longjmp) */
    if (setjmp(P->jb) == 0) {
        switch (op) {
            case 0: run_trim_fixed(P, a, na, (size_t)x, (size_t)y); break;
            case 1: run_check(P, a, na); break;
            case 2: run_statistics(P, a, na); break;
            case 3: run_interleave(P, a, na, b, nb); break;
            case 4: run_deinterleave(P, o2, a, na); break;
            case 5: run_dual_umi(P, a, na, (size_t)x); break;
        }
    }
    finish(P, R);
    free(P);
    if (out2) {
        *out2 = o2->p;
        *out2_n = o2->n;
    } else {
        free(o2->p);
    }
    free(o2);
    return R->exit_code;
}

demux_t *orc_demux(const uint8_t *sheet, size_t ns, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                   int has_r2, const uint8_t *i1, size_t ni1, int has_i1, const uint8_t *i2, size_t ni2, int has_i2,
                   uint64_t dry_run) {
    demux_t *D = (demux_t *)calloc(1, sizeof(demux_t));
    if (setjmp(D->P.jb) == 0) run_demux(D, sheet, ns, r1, n1, r2, n2, has_r2, i1, ni1, has_i1, i2, ni2, has_i2, dry_run);
    return D;
}
int orc_demux_exit_code(const demux_t *D) { return D->P.exit_code; }
size_t orc_demux_n_samples(const demux_t *D) { return D->n_samples; }
int orc_demux_paired(const demux_t *D) { return D->paired; }
int orc_demux_outputs_created(const demux_t *D) { return D->outputs_created; }
uint64_t orc_demux_total(const demux_t *D) { return D->total_reads; }
uint64_t orc_demux_identified(const demux_t *D) { return D->identified_reads; }
uint64_t orc_demux_sample_count(const demux_t *D, size_t s) { return D->samples[s].total_reads; }
const uint8_t *orc_demux_sample_name(const demux_t *D, size_t s, size_t *n) { *n = D->samples[s].name.n; return D->samples[s].name.p; }
const uint8_t *orc_demux_sample_out(const demux_t *D, size_t s, int mate, size_t *n) { *n = D->samples[s].output[mate].n; return D->samples[s].output[mate].p; }
const uint8_t *orc_demux_stdout(const demux_t *D, size_t *n) { *n = D->P.out.n; return D->P.out.p; }
const uint8_t *orc_demux_stderr(const demux_t *D, size_t *n) { *n = D->P.err.n; return D->P.err.p; }
void orc_demux_free(demux_t *D) {
    for (size_t s = 0; s < D->n_samples; s++) {
        free(D->samples[s].name.p); free(D->samples[s].barcode.p);
        free(D->samples[s].output[0].p); free(D->samples[s].output[1].p);
    }
    free(D->samples); free(D->P.out.p); free(D->P.err.p);
    free(D);
}

/* ------------------------------------------------------------------ CLI (`fasta_oracle`) */
#ifdef ORACLE_MAIN
#include <unistd.h>
#include <sys/wait.h>

/* FileReader::new, common.rs:88-104: "-" = stdin, *.gz through `gunzip -c`, else plain file. */
static uint8_t *slurp_fd(FILE *f, size_t *n) {
    obuf b = {0};
    uint8_t tmp[1 << 16];
    size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) ob_put(&b, tmp, k);
    *n = b.n;
    return b.p ? b.p : (uint8_t *)calloc(1, 1);
}
static uint8_t *open_input(const char *path, size_t *n) {
    if (strcmp(path, "-") == 0) return slurp_fd(stdin, n);
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "ERROR: Cannot open file %s for reading.\n", path);
        exit(255);
    }
    size_t L = strlen(path);
    if (L >= 3 && strcmp(path + L - 3, ".gz") == 0) {
        fclose(f);
        char cmd[4096];
        snprintf(cmd, sizeof cmd, "gunzip -c < '%s'", path);
        FILE *p = popen(cmd, "r");
        if (!p) { fprintf(stderr, "ERROR: Cannot start gunzip process.\n"); exit(255); }
        uint8_t *d = slurp_fd(p, n);
        pclose(p);
        return d;
    }
    uint8_t *d = slurp_fd(f, n);
    fclose(f);
    return d;
}
/* GzipWriter, common.rs:49-81: File::create(path) as stdout of `gzip -c` (or pigz). */
static void gzip_write(const char *path, const uint8_t *p, size_t n, int parallel) {
    FILE *chk = fopen(path, "wb");
    if (!chk) { fprintf(stderr, "ERROR: Cannot open file %s for writing.\n", path); exit(255); }
    fclose(chk);
    char cmd[4096];
    snprintf(cmd, sizeof cmd, "%s -c > '%s'", parallel ? "pigz" : "gzip", path);
    FILE *g = popen(cmd, "w");
    if (!g) { fprintf(stderr, "ERROR: Cannot start gzip process.\n"); exit(255); }
    fwrite(p, 1, n, g);
    pclose(g);
}

static const char *TOP_USAGE =
    "\nUsage:\n  fasta trim by quality <fastq_file> <min_baseq>\n  fasta mask by quality <fastq_file> <min_baseq>\n"
    "  fasta add barcode <fastq_file> <barcode_file>\n"
    "  fasta demultiplex [--parallel] [--index1=FASTQ] [--index2=FASTQ] [--dry-run=N] <sample_sheet> <fastq_1> [<fastq_2>]\n";

static int parse_u8(const char *s, unsigned *v) {
    if (!*s) return 0;
    const char *p = s;
    if (*p == '+') p++; /* Rust u8::from_str accepts a leading '+' */
    if (!*p) return 0;
    unsigned long x = 0;
    for (; *p; p++) {
        if (*p < '0' || *p > '9') return 0;
        x = x * 10 + (unsigned long)(*p - '0');
        if (x > 255) return 0;
    }
    *v = (unsigned)x;
    return 1;
}

int main(int argc, char **argv) {
    orc_result R = {0};
    if (argc >= 4 && !strcmp(argv[1], "trim") && !strcmp(argv[2], "by") && !strcmp(argv[3], "quality")) {
        if (argc != 6) { fprintf(stderr, "ERROR: Invalid arguments.\n\nUsage:\n  fasta trim by quality <fastq_file> <min_baseq>\n\n"); return 255; }
        size_t n; uint8_t *in = open_input(argv[4], &n);
        unsigned q;
        if (!parse_u8(argv[5], &q)) { fprintf(stderr, "thread 'main' panicked: min_baseq parse\n"); return 101; }
        orc_trim_by_quality(in, n, q, &R);
    } else if (argc >= 4 && !strcmp(argv[1], "mask") && !strcmp(argv[2], "by") && !strcmp(argv[3], "quality")) {
        if (argc != 6) { fprintf(stderr, "ERROR: Invalid arguments.\n\nUsage:\n  fasta mask by quality <fastq_file> <min_baseq>\n\n"); return 255; }
        size_t n; uint8_t *in = open_input(argv[4], &n);
        unsigned q;
        if (!parse_u8(argv[5], &q)) { fprintf(stderr, "thread 'main' panicked: min_baseq parse\n"); return 101; }
        orc_mask_by_quality(in, n, q, &R);
    } else if (argc >= 3 && !strcmp(argv[1], "add") && !strcmp(argv[2], "barcode")) {
        if (argc != 5) { fprintf(stderr, "ERROR: Invalid arguments.\n\nUsage:\n  fasta add barcode <fastq_file> <barcode_file>\n\n"); return 255; }
        size_t n, nb; uint8_t *fq = open_input(argv[3], &n); uint8_t *bc = open_input(argv[4], &nb);
        orc_add_barcode(fq, n, bc, nb, &R);
    } else if (argc >= 2 && (!strcmp(argv[1], "trim") || !strcmp(argv[1], "check") || !strcmp(argv[1], "statistics") ||
                             !strcmp(argv[1], "interleave") || !strcmp(argv[1], "deinterleave") ||
                             (argc >= 4 && !strcmp(argv[1], "extract") && !strcmp(argv[2], "dual") && !strcmp(argv[3], "umi")))) {
        /* fasta trim [--first=N] [--last=N] <f> | check <f> | statistics <f> | interleave <f1> <f2> |
         * deinterleave <f> <prefix> | extract dual umi [--first-bases=N] <f> */
        const int is_umi = !strcmp(argv[1], "extract");
        const int op = !strcmp(argv[1], "trim") ? 0 : !strcmp(argv[1], "check") ? 1 : !strcmp(argv[1], "statistics") ? 2
                       : !strcmp(argv[1], "interleave") ? 3 : !strcmp(argv[1], "deinterleave") ? 4 : 5;
        const char *pos[4]; int npos = 0; const char *first = "0", *last = "0", *fb = "0";
        for (int a = is_umi ? 4 : 2; a < argc; a++) {
            if (op == 0 && !strncmp(argv[a], "--first=", 8)) first = argv[a] + 8;
            else if (op == 0 && !strncmp(argv[a], "--last=", 7)) last = argv[a] + 7;
            else if (op == 5 && !strncmp(argv[a], "--first-bases=", 14)) fb = argv[a] + 14;
            else if (argv[a][0] == '-' && argv[a][1]) npos = 99;
            else if (npos < 4) pos[npos++] = argv[a];
        }
        const int want = (op == 3 || op == 4) ? 2 : 1;
        if (npos != want) { fprintf(stderr, "ERROR: Invalid arguments.\n"); return 255; }
        size_t na, nb = 0; uint8_t *a = open_input(pos[0], &na), *b = NULL;
        if (op == 3) b = open_input(pos[1], &nb);
        uint64_t x = 0, y = 0;
        const char *vals[2] = {op == 5 ? fb : first, last};
        const char *names[2] = {op == 5 ? "--first-bases=N" : "--first=N", "--last=N"};
        for (int k = 0; k < (op == 0 ? 2 : op == 5 ? 1 : 0); k++) { /* usize::from_str: optional '+', digits */
            const char *q = vals[k]; if (*q == '+') q++;
            char *e; unsigned long long v = strtoull(q, &e, 10);
            if (!*q || *e || q[0] < '0' || q[0] > '9') { fprintf(stderr, "ERROR: N must be a non-negative integer in %s.\n", names[k]); return 255; }
            if (k == 0) x = v; else y = v;
        }
        char p1[4096], p2[4096];
        if (op == 4) { /* GzipWriter::with_method creates the files first (fasta_deinterleave.rs:17-20) */
            snprintf(p1, sizeof p1, "%s_1.fq.gz", pos[1]);
            snprintf(p2, sizeof p2, "%s_2.fq.gz", pos[1]);
            FILE *c1 = fopen(p1, "wb"); if (!c1) { fprintf(stderr, "ERROR: Cannot open file %s for writing.\n", p1); return 255; } fclose(c1);
            FILE *c2 = fopen(p2, "wb"); if (!c2) { fprintf(stderr, "ERROR: Cannot open file %s for writing.\n", p2); return 255; } fclose(c2);
        }
        uint8_t *o2 = NULL; size_t o2n = 0;
        orc_next(op, a, na, b, nb, x, y, &R, &o2, &o2n);
        if (op == 4) {
            gzip_write(p1, R.out, R.out_n, 0);
            gzip_write(p2, o2, o2n, 0);
            R.out_n = 0;
        }
    } else if (argc >= 2 && !strcmp(argv[1], "demultiplex")) {
        const char *i1 = NULL, *i2 = NULL, *pos[3]; int npos = 0, parallel = 0; uint64_t dry = 0;
        for (int a = 2; a < argc; a++) {
            if (!strcmp(argv[a], "--parallel")) parallel = 1;
            else if (!strncmp(argv[a], "--index1=", 9)) i1 = argv[a] + 9;
            else if (!strncmp(argv[a], "--index2=", 9)) i2 = argv[a] + 9;
            else if (!strncmp(argv[a], "--dry-run=", 10)) {
                char *e; dry = strtoull(argv[a] + 10, &e, 10);
                if (*e || dry == 0) { fprintf(stderr, "ERROR: In --dry-run=N, N must be 64-bit positive integer.\n"); return 255; }
            } else if (npos < 3) pos[npos++] = argv[a];
            else npos = 99;
        }
        if (npos < 2 || npos > 3) { fprintf(stderr, "ERROR: Invalid arguments.\n"); return 255; }
        size_t n1, n2 = 0, ni1 = 0, ni2 = 0, ns;
        uint8_t *r1 = open_input(pos[1], &n1), *r2 = npos == 3 ? open_input(pos[2], &n2) : NULL;
        uint8_t *x1 = i1 ? open_input(i1, &ni1) : NULL, *x2 = i2 ? open_input(i2, &ni2) : NULL;
        uint8_t *sheet = open_input(pos[0], &ns);
        demux_t *D = orc_demux(sheet, ns, r1, n1, r2, n2, r2 != NULL, x1, ni1, x1 != NULL, x2, ni2, x2 != NULL, dry);
        if (D->outputs_created)
            for (size_t s = 0; s < D->n_samples; s++) {
                char path[4096];
                if (D->paired) {
                    snprintf(path, sizeof path, "%.*s_1.fq.gz", (int)D->samples[s].name.n, D->samples[s].name.p);
                    gzip_write(path, D->samples[s].output[0].p, D->samples[s].output[0].n, parallel);
                    snprintf(path, sizeof path, "%.*s_2.fq.gz", (int)D->samples[s].name.n, D->samples[s].name.p);
                    gzip_write(path, D->samples[s].output[1].p, D->samples[s].output[1].n, parallel);
                } else {
                    snprintf(path, sizeof path, "%.*s.fq.gz", (int)D->samples[s].name.n, D->samples[s].name.p);
                    gzip_write(path, D->samples[s].output[0].p, D->samples[s].output[0].n, parallel);
                }
            }
        fwrite(D->P.out.p, 1, D->P.out.n, stdout);
        fwrite(D->P.err.p, 1, D->P.err.n, stderr);
        return D->P.exit_code;
    } else {
        fprintf(stderr, "%s\n", TOP_USAGE); /* fasta_main.rs:79-81: usage on stderr, status 0 */
        return 0;
    }
    fwrite(R.out, 1, R.out_n, stdout);
    fwrite(R.err, 1, R.err_n, stderr);
    return R.exit_code;
}
#endif
