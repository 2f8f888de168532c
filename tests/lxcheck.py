"""Shared by the CPU and GPU lx(1) tests: generate a C lexer from a .lx spec with a given lx binary,
compile it with a tiny driver, tokenise a text and return the token stream."""
import os
import subprocess

DRIVER = r"""
#include <stdio.h>
#include <stdlib.h>
#include "lexer.h"
int main(void) {
	struct lx lx; struct lx_dynbuf buf; enum lx_token t;
	lx_init(&lx);
	lx.lgetc = lx_fgetc; lx.getc_opaque = stdin;
	buf.a = NULL; buf.len = 0;
	lx.buf_opaque = &buf; lx.push = lx_dynpush; lx.clear = lx_dynclear; lx.free = lx_dynfree;
	do {
		t = lx_next(&lx);
		printf("%u:%u %s [%s]\n", lx.start.line, lx.start.col, lx_name(t), buf.a != NULL ? buf.a : "");
	} while (t != TOK_EOF && t != TOK_ERROR && t != TOK_UNKNOWN);
	return 0;
}
"""


def token_stream(lx_binary: str, spec_path: str, text: bytes, workdir, concurrency: int = 1, env=None) -> bytes:
    # concurrency stays 1: with -C > 1 the REFERENCE's own lx races on an unsynchronised global table
    # (src/lx/ast.c:151-182, ast_setendmapping's mapping_count / realloc'd array -- ThreadSanitizer
    # reports it with the unmodified reference, and lx_ref -C 8 under load occasionally prints
    # nothing), so its output is not a stable yardstick.  The shim's own thread safety is covered by
    # shim_threads (TSAN-clean) and tests/test_gpu_threads.py.
    os.makedirs(workdir, exist_ok=True)
    for lang, name in (("h", "lexer.h"), ("c", "lexer.c")):
        p = subprocess.run([lx_binary, "-C", str(concurrency), "-l", lang, "-b", "dyn", "-g", "fgetc"], stdin=open(spec_path),
                           capture_output=True, timeout=300, env=env)
        assert p.returncode == 0, (lx_binary, lang, p.stderr[-500:])
        with open(os.path.join(workdir, name), "wb") as f:
            f.write(p.stdout)
    with open(os.path.join(workdir, "driver.c"), "w") as f:
        f.write(DRIVER)
    exe = os.path.join(workdir, "tokenise")
    subprocess.run(["gcc", "-std=c99", "-O1", "-w", "-DLX_HEADER=\"lexer.h\"", "-I", workdir, "-o", exe,
                    os.path.join(workdir, "driver.c"), os.path.join(workdir, "lexer.c")], check=True, timeout=300)
    p = subprocess.run([exe], input=text, capture_output=True, timeout=60)
    assert p.returncode == 0, p.stderr[-500:]
    return p.stdout


# (spec relative to the reference tree, sample text) -- the reference's own lexer specifications
SPECS = [
    ("src/libfsm/lexer.lx", b"0 -> 1 'a'; 1 -> 2 \"\\x41\"; # comment\n2 -> 2 ?;\nstart: 0;\nend: 2 = [1, 2];\n"),
    ("src/lx/lexer.lx", b"/[a-z]+/ -> $ident;\n'\"' .. '\"' -> $str { /./ -> $char; }\n# c\n"),
    ("src/libre/dialect/native/lexer.lx", b"ab*(c|d)+[a-z]{2,3}\\n.?"),
    ("src/libre/dialect/glob/lexer.lx", b"*.t?t[abc]"),
]

SAMPLE_SPEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sample.lx")
SAMPLE_TEXT = (b"foo = bar_1 -> 0x1F; // comment\n"
               b"x == 3.14 - 42 (\"a \\\"quoted\\\" \\\\ string\") y2;\n")
