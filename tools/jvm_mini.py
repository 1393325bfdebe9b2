"""A miniature JVM bytecode interpreter -- enough to EXECUTE the numeric leaf methods of the reference's own binaries
(output/MVTopicModel-1.0-SNAPSHOT.jar, output/lib/mallet-2.0.8.jar) in a container that has no JVM, so that the oracle's
restatements can be pinned against outputs of the reference itself (tests/golden/make_reference_vectors.py).

Scope: static and instance methods over ints / longs / floats / doubles, primitive arrays, plain objects with fields, calls into
other classes of the same jars, and a handful of java.lang / java.util natives (Math, Arrays.fill / copyOf, Object.<init>,
exception constructors).  No threads, no strings beyond constants, no exception tables (a Java throw becomes JavaThrow).
Java semantics kept: 32/64-bit wrap-around integer arithmetic, truncating division, shift masking, float rounding to binary32,
dcmpl/dcmpg NaN rules, d2i / d2l saturation.
"""
import math
import struct
import sys
import zipfile

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import jclass  # noqa: E402


class JavaThrow(Exception):
    pass


def i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def i64(x):
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x & (1 << 63) else x


def f32(x):
    try:
        return struct.unpack("f", struct.pack("f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _idiv(a, b):
    if b == 0:
        raise JavaThrow("java/lang/ArithmeticException")
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _d2int(x, bits):
    if x != x:
        return 0
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    if x <= lo:
        return lo
    if x >= hi:
        return hi
    return int(x)


def _fdiv(a, b):
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _frem(a, b):
    if b == 0.0 or a != a or b != b or math.isinf(a):
        return math.nan
    return math.fmod(a, b)


class JObject:
    def __init__(self, cls):
        self.cls, self.fields = cls, {}


NATIVES = {
    "java/lang/Math.log:(D)D": math.log, "java/lang/Math.exp:(D)D": math.exp, "java/lang/Math.sqrt:(D)D": math.sqrt,
    "java/lang/Math.pow:(DD)D": math.pow, "java/lang/Math.floor:(D)D": math.floor, "java/lang/Math.ceil:(D)D": math.ceil,
    "java/lang/Math.abs:(D)D": abs, "java/lang/Math.abs:(I)I": lambda a: i32(abs(a)), "java/lang/Math.max:(II)I": max,
    "java/lang/Math.min:(II)I": min, "java/lang/Math.max:(DD)D": max, "java/lang/Math.min:(DD)D": min,
    "java/lang/Math.round:(D)J": lambda a: _d2int(math.floor(a + 0.5), 64),
    "java/lang/Double.isNaN:(D)Z": lambda a: int(a != a),
    "java/lang/Double.isInfinite:(D)Z": lambda a: int(math.isinf(a)),
    "java/lang/Class.desiredAssertionStatus:()Z": lambda: 0,
}


def _jlog(x):
    if x != x or x < 0:
        return math.nan
    if x == 0:
        return -math.inf
    return math.log(x)


NATIVES["java/lang/Math.log:(D)D"] = _jlog


class MiniJVM:
    def __init__(self, jars):
        self.zips = [zipfile.ZipFile(j) for j in jars]
        self.classes, self.statics, self.inited = {}, {}, set()
        self.steps = 0
        # host shims for library classes: "class.method:descriptor" -> f(frame_locals, receiver, args, call_site_pc).  They see
        # the calling frame's locals, so a shim can e.g. key a random draw on the loop variables of the method that asked for it
        self.shims = {}
        # debugging aid: {(len(code), pc): f(frame_locals, operand_stack)} called before the instruction at pc executes
        self.probes = {}
        self.strict_fields = False      # True: reading a field nobody has written raises instead of yielding the JVM default

    def load(self, name):
        if name not in self.classes:
            for z in self.zips:
                try:
                    self.classes[name] = jclass.ClassFile(z.read(name + ".class"))
                    break
                except KeyError:
                    continue
            else:
                raise KeyError(name)
        cf = self.classes[name]
        if name not in self.inited:
            self.inited.add(name)
            for n, d, code in cf.methods:
                if n == "<clinit>" and code:
                    try:
                        self.run(cf, code, [])
                    except (NotImplementedError, KeyError):
                        pass        # loggers and other non-numeric statics: whatever was assigned before stays
        return cf

    def find(self, cls, name, desc):
        cf = self.load(cls)
        for n, d, code in cf.methods:
            if n == name and d == desc:
                return cf, code
        raise KeyError(f"{cls}.{name}{desc}")

    @staticmethod
    def nargs(desc):
        """slot-aware argument kinds of a descriptor: list of 1 (one slot) / 2 (long, double)"""
        out, i = [], 1
        while desc[i] != ")":
            c = desc[i]
            if c in "JD":
                out.append(2); i += 1
            elif c == "L":
                out.append(1); i = desc.index(";", i) + 1
            elif c == "[":
                while desc[i] == "[":
                    i += 1
                i = desc.index(";", i) + 1 if desc[i] == "L" else i + 1
                out.append(1)
            else:
                out.append(1); i += 1
        return out

    def call(self, cls, name, desc, args):
        """args: Python values (ints, floats, lists for arrays, JObject); instance methods take the receiver first."""
        cf, code = self.find(cls, name, desc)
        kinds = self.nargs(desc)
        static = any(n == name and d == desc and True for n, d, _ in cf.methods) and len(args) == len(kinds)
        loc = []
        vals = list(args)
        if not static:
            loc.append(vals.pop(0))
        for v, k in zip(vals, kinds):
            loc.append(v)
            if k == 2:
                loc.append(None)
        return self.run(cf, code, loc)

    def new(self, cls, desc="()V", args=()):
        self.load(cls)
        o = JObject(cls)
        self.call(cls, "<init>", desc, [o] + list(args))
        return o

    def run(self, cf, code, loc):
        try:
            return self._run(cf, code, loc)
        except JavaThrow:
            raise
        except Exception as e:          # annotate host-side failures with the bytecode location (innermost frame only)
            if not getattr(e, "_jvm_where", None):
                e._jvm_where = (cf.this, self._pc)
                e.args = (f"{e.args[0] if e.args else ''} [at {cf.this} pc {self._pc}]",) + tuple(e.args[1:])
            raise

    def _run(self, cf, code, loc):
        loc = list(loc) + [None] * 64
        st, pc, u = [], 0, struct.unpack_from
        cp = cf.cp

        def ref(i):
            c = cp[i]
            nat = cp[c[2]]
            return cf.cname(c[1]), cf.utf(nat[1]), cf.utf(nat[2])
        ncode = len(code)
        while True:
            self.steps += 1
            self._pc = pc
            if self.probes and (ncode, pc) in self.probes:
                self.probes[(ncode, pc)](loc, st)
            op = code[pc]
            if op == 0: pc += 1
            elif op == 1: st.append(None); pc += 1
            elif 2 <= op <= 8: st.append(op - 3); pc += 1
            elif op in (9, 10): st.append(op - 9); pc += 1
            elif 11 <= op <= 13: st.append(float(op - 11)); pc += 1
            elif op in (14, 15): st.append(float(op - 14)); pc += 1
            elif op == 16: st.append(u(">b", code, pc + 1)[0]); pc += 2
            elif op == 17: st.append(u(">h", code, pc + 1)[0]); pc += 3
            elif op in (18, 19, 20):
                idx = code[pc + 1] if op == 18 else u(">H", code, pc + 1)[0]
                c = cp[idx]
                st.append(cf.utf(c[1]) if c[0] == "string" else c[1]); pc += 2 if op == 18 else 3
            elif 21 <= op <= 25: st.append(loc[code[pc + 1]]); pc += 2
            elif 26 <= op <= 45: st.append(loc[(op - 26) % 4]); pc += 1
            elif 46 <= op <= 53:
                i = st.pop(); a = st.pop()
                if a is None: raise JavaThrow("java/lang/NullPointerException")
                if not 0 <= i < len(a): raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException")
                st.append(a[i]); pc += 1
            elif 54 <= op <= 58: loc[code[pc + 1]] = st.pop(); pc += 2
            elif 59 <= op <= 78: loc[(op - 59) % 4] = st.pop(); pc += 1
            elif 79 <= op <= 86:
                v = st.pop(); i = st.pop(); a = st.pop()
                if not 0 <= i < len(a): raise JavaThrow("java/lang/ArrayIndexOutOfBoundsException")
                if op == 81: v = f32(v)
                elif op == 84: v = ((v & 0xFF) ^ 0x80) - 0x80
                elif op == 85: v &= 0xFFFF
                elif op == 86: v = ((v & 0xFFFF) ^ 0x8000) - 0x8000
                a[i] = v; pc += 1
            elif op == 87: st.pop(); pc += 1
            elif op == 88:
                st.pop(); pc += 1          # values are one Python slot whatever their category: pop2 of a long/double pops one
            elif op == 89: st.append(st[-1]); pc += 1
            elif op == 90: st.insert(-2, st[-1]); pc += 1
            elif op == 92:
                # dup2: category-2 value = one Python slot; category-1 pair = two.  Decide by Python type of the top value
                if isinstance(st[-1], float) or getattr(self, "_top_is_wide", False): st.append(st[-1])
                else: st.extend(st[-2:])
                pc += 1
            elif op == 95: st[-1], st[-2] = st[-2], st[-1]; pc += 1
            elif 96 <= op <= 115:
                b = st.pop(); a = st.pop(); k = (op - 96) % 4; g = (op - 96) // 4
                if k == 0 or k == 1:
                    w = i32 if k == 0 else i64
                    r = (a + b, a - b, a * b, None, None)[g] if g < 3 else (_idiv(a, b) if g == 3 else a - _idiv(a, b) * b)
                    st.append(w(r))
                else:
                    r = (a + b, a - b, a * b)[g] if g < 3 else (_fdiv(a, b) if g == 3 else _frem(a, b))
                    st.append(f32(r) if k == 2 else r)
                pc += 1
            elif 116 <= op <= 119:
                a = st.pop(); k = op - 116
                st.append(i32(-a) if k == 0 else i64(-a) if k == 1 else -a); pc += 1
            elif 120 <= op <= 125:
                b = st.pop(); a = st.pop(); long_ = (op - 120) % 2; g = (op - 120) // 2
                bits = 64 if long_ else 32; s = b & (bits - 1); w = i64 if long_ else i32
                if g == 0: r = a << s
                elif g == 1: r = a >> s
                else: r = (a & ((1 << bits) - 1)) >> s
                st.append(w(r)); pc += 1
            elif 126 <= op <= 131:
                b = st.pop(); a = st.pop(); g = (op - 126) // 2
                st.append((a & b, a | b, a ^ b)[g]); pc += 1
            elif op == 132:
                loc[code[pc + 1]] = i32(loc[code[pc + 1]] + u(">b", code, pc + 2)[0]); pc += 3
            elif 133 <= op <= 147:
                a = st.pop()
                r = {133: lambda: a, 134: lambda: f32(float(a)), 135: lambda: float(a), 136: lambda: i32(a), 137: lambda: f32(float(a)),
                     138: lambda: float(a), 139: lambda: _d2int(a, 32), 140: lambda: _d2int(a, 64), 141: lambda: float(a),
                     142: lambda: _d2int(a, 32), 143: lambda: _d2int(a, 64), 144: lambda: f32(a),
                     145: lambda: ((a & 0xFF) ^ 0x80) - 0x80, 146: lambda: a & 0xFFFF, 147: lambda: ((a & 0xFFFF) ^ 0x8000) - 0x8000}[op]()
                st.append(r); pc += 1
            elif op == 148:
                b = st.pop(); a = st.pop(); st.append((a > b) - (a < b)); pc += 1
            elif 149 <= op <= 152:
                b = st.pop(); a = st.pop()
                if a != a or b != b: st.append(-1 if op in (149, 151) else 1)
                else: st.append((a > b) - (a < b))
                pc += 1
            elif 153 <= op <= 158:
                a = st.pop(); t = (a == 0, a != 0, a < 0, a >= 0, a > 0, a <= 0)[op - 153]
                pc = pc + u(">h", code, pc + 1)[0] if t else pc + 3
            elif 159 <= op <= 164:
                b = st.pop(); a = st.pop(); t = (a == b, a != b, a < b, a >= b, a > b, a <= b)[op - 159]
                pc = pc + u(">h", code, pc + 1)[0] if t else pc + 3
            elif op in (165, 166):
                b = st.pop(); a = st.pop(); t = (a is b) if op == 165 else (a is not b)
                pc = pc + u(">h", code, pc + 1)[0] if t else pc + 3
            elif op == 167: pc += u(">h", code, pc + 1)[0]
            elif 172 <= op <= 176: return st.pop()
            elif op == 177: return None
            elif op == 178:
                c, n, d = ref(u(">H", code, pc + 1)[0])
                if n == "$assertionsDisabled": st.append(1)
                elif c.startswith("java/"): st.append(JObject(c + "." + n))
                else:
                    self.load(c); st.append(self.statics.get((c, n), 0.0 if d in "DF" else 0 if d in "IJSBCZ" else None))
                pc += 3
            elif op == 179:
                c, n, d = ref(u(">H", code, pc + 1)[0]); self.statics[(c, n)] = st.pop(); pc += 3
            elif op == 180:
                c, n, d = ref(u(">H", code, pc + 1)[0]); o = st.pop()
                if self.strict_fields and n not in o.fields:
                    raise KeyError(f"field {c}.{n}:{d} was never set")
                st.append(o.fields.get(n, 0.0 if d in "DF" else 0 if d in "IJSBCZ" else None)); pc += 3
            elif op == 181:
                c, n, d = ref(u(">H", code, pc + 1)[0]); v = st.pop(); o = st.pop(); o.fields[n] = v; pc += 3
            elif op in (182, 183, 184, 185):
                c, n, d = ref(u(">H", code, pc + 1)[0])
                kinds = self.nargs(d)
                args = [st.pop() for _ in kinds][::-1]
                recv = st.pop() if op != 184 else None
                key = f"{c}.{n}:{d}"
                if key in self.shims:
                    r = self.shims[key](loc, recv, args, pc)
                elif key in NATIVES:
                    r = NATIVES[key](*args)
                elif c == "java/lang/Object" and n == "<init>":
                    r = None
                elif c == "java/util/Arrays" and n == "fill":
                    for i in range(len(args[0])): args[0][i] = args[1]
                    r = None
                elif c == "java/util/Arrays" and n == "copyOf":
                    r = (list(args[0]) + [0.0 if d.startswith("([D") else 0] * args[1])[:args[1]]
                elif c.startswith("java/") and n == "<init>":
                    r = None                                  # exception / builder constructors: nothing to do
                elif c.startswith("java/"):
                    raise NotImplementedError(key)
                else:
                    target = recv.cls if (op in (182, 185) and isinstance(recv, JObject)) else c
                    tcf, tcode = self.find(target, n, d)
                    l2 = [] if op == 184 else [recv]
                    for v, k in zip(args, kinds):
                        l2.append(v)
                        if k == 2: l2.append(None)
                    r = self.run(tcf, tcode, l2)
                if not d.endswith(")V"): st.append(r)
                pc += 5 if op == 185 else 3
            elif op == 187:
                st.append(JObject(cf.cname(u(">H", code, pc + 1)[0]))); pc += 3
            elif op == 188:
                n = st.pop(); t = code[pc + 1]
                if n < 0: raise JavaThrow("java/lang/NegativeArraySizeException")
                st.append([0.0] * n if t in (6, 7) else [0] * n); pc += 2
            elif op == 189:
                n = st.pop(); st.append([None] * n); pc += 3
            elif op == 197:
                dims = code[pc + 3]
                sizes = [st.pop() for _ in range(dims)][::-1]
                desc = cf.cname(u(">H", code, pc + 1)[0])
                leaf0 = 0.0 if desc.lstrip("[")[:1] in "DF" else (0 if desc.lstrip("[")[:1] in "IJSBCZ" else None)

                def mk(level):
                    if level == len(sizes) - 1:
                        return [leaf0 if desc.count("[") == len(sizes) else None] * sizes[level]
                    return [mk(level + 1) for _ in range(sizes[level])]
                st.append(mk(0)); pc += 4
            elif op == 190:
                a = st.pop()
                if a is None: raise JavaThrow("java/lang/NullPointerException")
                st.append(len(a)); pc += 1
            elif op == 191:
                o = st.pop(); raise JavaThrow(o.cls if isinstance(o, JObject) else str(o))
            elif op == 192: pc += 3
            elif op == 193:            # instanceof: host objects are trusted to be what the bytecode expects; null is an instance of nothing
                o = st.pop(); st.append(0 if o is None else 1); pc += 3
            elif op in (198, 199):
                a = st.pop(); t = (a is None) if op == 198 else (a is not None)
                pc = pc + u(">h", code, pc + 1)[0] if t else pc + 3
            else:
                raise NotImplementedError(f"opcode {op} ({jclass.OPS.get(op, ('?',))[0]}) at pc {pc}")


if __name__ == "__main__":
    vm = MiniJVM(["/root/reference/output/lib/mallet-2.0.8.jar", "/root/reference/output/MVTopicModel-1.0-SNAPSHOT.jar"])
    print("logGammaStirling(0.1) =", repr(vm.call("cc/mallet/types/Dirichlet", "logGammaStirling", "(D)D", [0.1])))
    t = vm.new("org/madgik/utils/FTree", "([D)V", [[1.0, 2.0, 3.0, 4.0]])
    print("FTree{1,2,3,4}.tree =", t.fields["tree"], "sample(0.4) =", vm.call("org/madgik/utils/FTree", "sample", "(D)I", [t, 0.4]))
