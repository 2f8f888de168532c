"""Shared by the CPU and GPU rx(1) tests: generate a C matcher from a pattern file with a given rx
binary (`rx -l c -k str`: int fsm_main(const char *s, unsigned *id)), compile it with a tiny driver
and return its verdicts for a list of strings."""
import os
import subprocess

DRIVER = r"""
#include <stdio.h>
#include <string.h>
int fsm_main(const char *s, unsigned *id);
int main(void) {
	char line[4096];
	while (fgets(line, sizeof line, stdin) != NULL) {
		unsigned id = 99999;
		size_t n = strlen(line);
		if (n > 0 && line[n - 1] == '\n') line[n - 1] = '\0';
		{ int r = fsm_main(line, &id); printf("%d %u\n", r, r ? id : 0u); }
	}
	return 0;
}
"""

# anchored, mutually exclusive patterns: rx rejects ambiguous sets (src/rx/main.c:568-)
PATTERNS = [r"^ERROR [0-9]+$", r"^user=[a-z]+$", r"^GET /[a-z]*( HTTP/1\.[01])?$", r"^(foo|bar)+$",
            r"^[0-9]+\.[0-9]+\.[0-9]+\.[0-9]+$", r"^x[^y]*y$"]
STRINGS = ["ERROR 42", "ERROR", "user=alice", "user=", "GET /", "GET /index HTTP/1.1", "GET /index HTTP/1.2",
           "foobarfoo", "fo", "10.0.0.1", "10.0.0", "xaaay", "xy", "xyy", "", "nothing"]


def verdicts(rx_binary: str, workdir, env=None) -> bytes:
    os.makedirs(workdir, exist_ok=True)
    pats = os.path.join(workdir, "patterns.txt")
    with open(pats, "w") as f:
        f.write("\n".join(PATTERNS) + "\n")
    p = subprocess.run([rx_binary, "-l", "c", "-k", "str", "-r", "pcre", pats], capture_output=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout, (rx_binary, p.stderr[-800:])
    with open(os.path.join(workdir, "matcher.c"), "wb") as f:
        f.write(p.stdout)
    with open(os.path.join(workdir, "driver.c"), "w") as f:
        f.write(DRIVER)
    exe = os.path.join(workdir, "match")
    subprocess.run(["gcc", "-std=c99", "-O1", "-w", "-o", exe, os.path.join(workdir, "driver.c"), os.path.join(workdir, "matcher.c")],
                   check=True, timeout=300)
    r = subprocess.run([exe], input=("\n".join(STRINGS) + "\n").encode(), capture_output=True, timeout=60)
    assert r.returncode == 0
    return r.stdout
