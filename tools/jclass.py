"""Minimal JVM class-file disassembler (constant pool + Code attributes), used to recover the arithmetic of binary-only
third-party classes next to the hot path (MALLET 2.0.8) when no JDK is available.

    python tools/jclass.py some.jar cc/mallet/types/IDSorter.class [method-name]

Prints every method (or the named one) as `pc: opcode operands`, constant-pool references resolved.  Reads, never executes.
"""
import struct
import sys
import zipfile

OPS = {}


def _def(names, start, fmt=""):
    for i, n in enumerate(names.split()):
        OPS[start + i] = (n, fmt)


_def("nop aconst_null iconst_m1 iconst_0 iconst_1 iconst_2 iconst_3 iconst_4 iconst_5 lconst_0 lconst_1 fconst_0 fconst_1 fconst_2 dconst_0 dconst_1", 0)
OPS[16] = ("bipush", "b"); OPS[17] = ("sipush", "h"); OPS[18] = ("ldc", "C"); OPS[19] = ("ldc_w", "W"); OPS[20] = ("ldc2_w", "W")
_def("iload lload fload dload aload", 21, "B")
_def("iload_0 iload_1 iload_2 iload_3 lload_0 lload_1 lload_2 lload_3 fload_0 fload_1 fload_2 fload_3 dload_0 dload_1 dload_2 dload_3 "
     "aload_0 aload_1 aload_2 aload_3 iaload laload faload daload aaload baload caload saload", 26)
_def("istore lstore fstore dstore astore", 54, "B")
_def("istore_0 istore_1 istore_2 istore_3 lstore_0 lstore_1 lstore_2 lstore_3 fstore_0 fstore_1 fstore_2 fstore_3 dstore_0 dstore_1 dstore_2 "
     "dstore_3 astore_0 astore_1 astore_2 astore_3 iastore lastore fastore dastore aastore bastore castore sastore pop pop2 dup dup_x1 dup_x2 "
     "dup2 dup2_x1 dup2_x2 swap iadd ladd fadd dadd isub lsub fsub dsub imul lmul fmul dmul idiv ldiv fdiv ddiv irem lrem frem drem ineg lneg "
     "fneg dneg ishl lshl ishr lshr iushr lushr iand land ior lor ixor lxor", 59)
OPS[132] = ("iinc", "Bb")
_def("i2l i2f i2d l2i l2f l2d f2i f2l f2d d2i d2l d2f i2b i2c i2s lcmp fcmpl fcmpg dcmpl dcmpg", 133)
_def("ifeq ifne iflt ifge ifgt ifle if_icmpeq if_icmpne if_icmplt if_icmpge if_icmpgt if_icmple if_acmpeq if_acmpne goto jsr", 153, "J")
OPS[169] = ("ret", "B"); OPS[170] = ("tableswitch", "T"); OPS[171] = ("lookupswitch", "L")
_def("ireturn lreturn freturn dreturn areturn return", 172)
_def("getstatic putstatic getfield putfield invokevirtual invokespecial invokestatic", 178, "W")
OPS[185] = ("invokeinterface", "Wxx"); OPS[186] = ("invokedynamic", "Wxx"); OPS[187] = ("new", "W"); OPS[188] = ("newarray", "B")
OPS[189] = ("anewarray", "W"); OPS[190] = ("arraylength", ""); OPS[191] = ("athrow", ""); OPS[192] = ("checkcast", "W")
OPS[193] = ("instanceof", "W"); OPS[194] = ("monitorenter", ""); OPS[195] = ("monitorexit", ""); OPS[196] = ("wide", "")
OPS[197] = ("multianewarray", "WB"); OPS[198] = ("ifnull", "J"); OPS[199] = ("ifnonnull", "J"); OPS[200] = ("goto_w", "I")


class ClassFile:
    def __init__(self, data):
        self.d, self.p = data, 0
        assert self.u4() == 0xCAFEBABE
        self.u2(); self.u2()
        n = self.u2()
        self.cp = [None] * n
        i = 1
        while i < n:
            t = self.u1()
            if t == 1:
                ln = self.u2(); self.cp[i] = ("utf8", self.d[self.p:self.p + ln].decode("utf-8", "replace")); self.p += ln
            elif t == 3: self.cp[i] = ("int", struct.unpack(">i", self.raw(4))[0])
            elif t == 4: self.cp[i] = ("float", struct.unpack(">f", self.raw(4))[0])
            elif t == 5: self.cp[i] = ("long", struct.unpack(">q", self.raw(8))[0]); i += 1
            elif t == 6: self.cp[i] = ("double", struct.unpack(">d", self.raw(8))[0]); i += 1
            elif t in (7, 8, 16, 19, 20): self.cp[i] = ({7: "class", 8: "string", 16: "mtype", 19: "module", 20: "package"}[t], self.u2())
            elif t in (9, 10, 11, 12): self.cp[i] = ({9: "field", 10: "method", 11: "imethod", 12: "nat"}[t], self.u2(), self.u2())
            elif t == 15: self.cp[i] = ("mhandle", self.u1(), self.u2())
            elif t in (17, 18): self.cp[i] = ("dyn", self.u2(), self.u2())
            else: raise ValueError(f"constant tag {t}")
            i += 1
        self.u2(); self.this = self.cname(self.u2()); self.u2()
        for _ in range(self.u2()): self.u2()
        self.fields = [self.member() for _ in range(self.u2())]
        self.methods = [self.member() for _ in range(self.u2())]

    def raw(self, n):
        b = self.d[self.p:self.p + n]; self.p += n; return b

    def u1(self): return self.raw(1)[0]
    def u2(self): return struct.unpack(">H", self.raw(2))[0]
    def u4(self): return struct.unpack(">I", self.raw(4))[0]
    def utf(self, i): return self.cp[i][1]
    def cname(self, i): return self.utf(self.cp[i][1])

    def ref(self, i):
        c = self.cp[i]
        if c is None: return f"#{i}"
        k = c[0]
        if k == "utf8": return repr(c[1])
        if k in ("int", "float", "long", "double"): return f"{k} {c[1]!r}"
        if k == "class": return self.utf(c[1])
        if k == "string": return "String " + repr(self.utf(c[1]))
        if k in ("field", "method", "imethod"):
            nat = self.cp[c[2]]
            return f"{self.cname(c[1])}.{self.utf(nat[1])}:{self.utf(nat[2])}"
        return str(c)

    def member(self):
        self.u2(); name = self.utf(self.u2()); desc = self.utf(self.u2())
        code = None
        for _ in range(self.u2()):
            an = self.utf(self.u2()); ln = self.u4(); body = self.raw(ln)
            if an == "Code":
                cl = struct.unpack(">I", body[4:8])[0]
                code = body[8:8 + cl]
        return name, desc, code

    def disasm(self, code):
        out, pc = [], 0
        while pc < len(code):
            op = code[pc]; name, fmt = OPS.get(op, (f"op{op}", "")); start = pc; pc += 1; args = []
            for f in fmt:
                if f == "b": args.append(str(struct.unpack(">b", code[pc:pc + 1])[0])); pc += 1
                elif f == "B": args.append(str(code[pc])); pc += 1
                elif f == "h": args.append(str(struct.unpack(">h", code[pc:pc + 2])[0])); pc += 2
                elif f == "C": args.append(self.ref(code[pc])); pc += 1
                elif f == "W": args.append(self.ref(struct.unpack(">H", code[pc:pc + 2])[0])); pc += 2
                elif f == "J": args.append("-> %d" % (start + struct.unpack(">h", code[pc:pc + 2])[0])); pc += 2
                elif f == "I": args.append("-> %d" % (start + struct.unpack(">i", code[pc:pc + 4])[0])); pc += 4
                elif f == "x": pc += 1
                elif f in "TL":
                    pc = (pc + 3) & ~3
                    dflt = struct.unpack(">i", code[pc:pc + 4])[0]; pc += 4
                    if f == "T":
                        lo, hi = struct.unpack(">ii", code[pc:pc + 8]); pc += 8
                        offs = struct.unpack(">%di" % (hi - lo + 1), code[pc:pc + 4 * (hi - lo + 1)]); pc += 4 * (hi - lo + 1)
                        args.append(" ".join(f"{lo + k}->{start + o}" for k, o in enumerate(offs)) + f" default->{start + dflt}")
                    else:
                        n = struct.unpack(">i", code[pc:pc + 4])[0]; pc += 4
                        prs = [struct.unpack(">ii", code[pc + 8 * k:pc + 8 * k + 8]) for k in range(n)]; pc += 8 * n
                        args.append(" ".join(f"{k}->{start + o}" for k, o in prs) + f" default->{start + dflt}")
            out.append(f"{start:5d}: {name} {' '.join(args)}")
        return out


def main():
    jar, member = sys.argv[1], sys.argv[2]
    only = sys.argv[3] if len(sys.argv) > 3 else None
    data = zipfile.ZipFile(jar).read(member) if jar.endswith(".jar") else open(jar, "rb").read()
    cf = ClassFile(data)
    print("class", cf.this)
    for name, desc, _ in cf.fields:
        print("  field", name, desc)
    for name, desc, code in cf.methods:
        if only and name != only:
            continue
        print(f"\nmethod {name}{desc}")
        if code:
            print("\n".join(cf.disasm(code)))


if __name__ == "__main__":
    main()
