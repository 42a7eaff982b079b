"""Build step of oracle/_ref (TEST INFRASTRUCTURE ONLY): cuts the MatMult row loops out of the
reference's own patch files WHERE THEY LIE under the reference tree and writes them, unmodified,
as include fragments into oracle/_ref/ (git-ignored; nothing of the reference enters the repo).

  matmult_original_loop.inc   old side (context and '-' lines) of src/openacc-step1/
                              MatMult_SeqAIJ.patch: the loop of PETSc 3.7.6's MatMult_SeqAIJ as the
                              reference's patch shows it (for ... { ... PetscSparseDensePlusDot ... })
  matmult_step3_host.inc      new side ('+' and context lines) of src/openacc-step3/
                              MatMult_SeqAIJ.patch from `PetscInt offset = 0;` to the end of the
                              device loop: the reference author's host loop + kernel loop
  matmult_step4_blocked.inc   the same region of src/openacc-step4/MatMult_SeqAIJ.patch: host loop,
                              the row-blocked kernel loops (983,040 rows per block) and the loop
                              over the remaining rows

usage: python extract_ref_loops.py <reference root> <output dir>
"""
import os
import re
import sys


def sides(patch_path):
    """(old, new): the hunks' old-side and new-side lines, in order."""
    old, new = [], []
    with open(patch_path) as f:
        for line in f.read().split("\n"):
            if line.startswith(("---", "+++", "@@")) or line == "":
                continue
            tag, body = line[0], line[1:]
            if tag in " -":
                old.append(body)
            if tag in " +":
                new.append(body)
    return old, new


def block_from(lines, start_pat, open_pat):
    """Lines from the first match of start_pat through the brace block opened at/after the first
    later match of open_pat."""
    s = next(i for i, l in enumerate(lines) if re.search(start_pat, l))
    o = next(i for i in range(s, len(lines)) if re.search(open_pat, lines[i]))
    depth, seen = 0, False
    for e in range(o, len(lines)):
        code = lines[e].split("//")[0]
        depth += code.count("{") - code.count("}")
        seen = seen or "{" in code
        if seen and depth == 0:
            return lines[s:e + 1]
    raise SystemExit("unterminated block")


def main(ref, out):
    os.makedirs(out, exist_ok=True)
    old1, _ = sides(os.path.join(ref, "src/openacc-step1/MatMult_SeqAIJ.patch"))
    loop = block_from(old1, r"for \(i=0; i<m; i\+\+\) \{", r"for \(i=0; i<m; i\+\+\) \{")
    _, new3 = sides(os.path.join(ref, "src/openacc-step3/MatMult_SeqAIJ.patch"))
    host = block_from(new3, r"PetscInt offset = 0;", r"for \(i=offset; i<m; i\+\+\) \{")
    _, new4 = sides(os.path.join(ref, "src/openacc-step4/MatMult_SeqAIJ.patch"))
    blocked = block_from(new4, r"PetscInt offset = 0;", r"for \(i=offset; i<m; i\+\+\) \{")
    for name, lines in (("matmult_original_loop.inc", loop), ("matmult_step3_host.inc", host),
                        ("matmult_step4_blocked.inc", blocked)):
        with open(os.path.join(out, name), "w") as f:
            f.write("\n".join(lines) + "\n")
        print(f"{name}: {len(lines)} lines")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
