"""Tiny, independent FlatBuffers reader/writer in pure Python (written from the format
specification, shares no code with csrc/wire/fb.h).  Used to cross-check the native codec:
whatever C++ encodes must decode here, and what this writer emits must decode in C++."""
import struct


# ----------------------------------------------------------------------------- reader
def _u16(b, o):
    return struct.unpack_from("<H", b, o)[0]


def _u32(b, o):
    return struct.unpack_from("<I", b, o)[0]


def _i32(b, o):
    return struct.unpack_from("<i", b, o)[0]


class Table:
    def __init__(self, buf, pos):
        self.buf, self.pos = buf, pos
        self.vt = pos - _i32(buf, pos)
        self.vt_len = _u16(buf, self.vt)

    def _field(self, slot):
        if slot + 2 > self.vt_len:
            return 0
        off = _u16(self.buf, self.vt + slot)
        return self.pos + off if off else 0

    def scalar(self, slot, fmt, default=0):
        f = self._field(slot)
        return struct.unpack_from("<" + fmt, self.buf, f)[0] if f else default

    def _indirect(self, slot):
        f = self._field(slot)
        return f + _u32(self.buf, f) if f else 0

    def string(self, slot):
        p = self._indirect(slot)
        if not p:
            return None
        n = _u32(self.buf, p)
        return bytes(self.buf[p + 4:p + 4 + n])

    def vector(self, slot, fmt):
        p = self._indirect(slot)
        if not p:
            return None
        n = _u32(self.buf, p)
        size = struct.calcsize("<" + fmt)
        return [struct.unpack_from("<" + fmt, self.buf, p + 4 + i * size) for i in range(n)]

    def string_vector(self, slot):
        p = self._indirect(slot)
        if not p:
            return None
        n = _u32(self.buf, p)
        out = []
        for i in range(n):
            loc = p + 4 + 4 * i
            s = loc + _u32(self.buf, loc)
            ln = _u32(self.buf, s)
            out.append(bytes(self.buf[s + 4:s + 4 + ln]))
        return out

    def table_vector(self, slot):
        p = self._indirect(slot)
        if not p:
            return None
        n = _u32(self.buf, p)
        out = []
        for i in range(n):
            loc = p + 4 + 4 * i
            out.append(Table(self.buf, loc + _u32(self.buf, loc)))
        return out


def root(buf):
    return Table(buf, _u32(buf, 0))


# ----------------------------------------------------------------------------- writer
def build_table(fields):
    """fields: list of (slot, kind, value) with kind in {'i8','i32','u32','u64','str',
    'vec_u64','vec_u8','vec_str','vec_struct16','vec_tab'}.  Returns bytes of a finished
    buffer.  Layout: [root uoffset][vtable][table][children...] — children AFTER the table
    so that every uoffset is positive."""
    # compute table inline layout
    inline = []  # (slot, fmt, size)
    for slot, kind, value in fields:
        if kind == "i8":
            inline.append((slot, "b", 1))
        elif kind in ("i32",):
            inline.append((slot, "i", 4))
        elif kind in ("u32",):
            inline.append((slot, "I", 4))
        elif kind == "u64":
            inline.append((slot, "Q", 8))
        else:
            inline.append((slot, "I", 4))  # uoffset
    # place fields after the 4-byte soffset, biggest alignment first
    order = sorted(range(len(inline)), key=lambda i: -inline[i][2])
    max_slot = max([s for s, _, _ in fields], default=2)
    vt_len = max_slot + 2
    vt_len += vt_len % 2
    # table must start 8-aligned relative to buffer start; vtable right before it
    buf = bytearray(4)
    while (len(buf) + vt_len) % 8:
        buf.append(0)
    vt_pos = len(buf)
    buf += bytes(vt_len)
    tab_pos = len(buf)
    buf += struct.pack("<i", tab_pos - vt_pos)
    offsets = {}
    for i in order:
        slot, fmt, size = inline[i]
        while (len(buf) - tab_pos) % size:
            buf.append(0)
        offsets[i] = len(buf) - tab_pos
        buf += bytes(size)
    tab_len = len(buf) - tab_pos
    struct.pack_into("<HH", buf, vt_pos, vt_len, tab_len)
    for i, (slot, _, _) in enumerate(inline):
        struct.pack_into("<H", buf, vt_pos + slot, offsets[i])
    struct.pack_into("<I", buf, 0, tab_pos)

    def patch(i, target):
        loc = tab_pos + offsets[i]
        struct.pack_into("<I", buf, loc, target - loc)

    def emit_string(s):
        while len(buf) % 4:
            buf.append(0)
        pos = len(buf)
        buf.extend(struct.pack("<I", len(s)) + s + b"\0")
        return pos

    for i, (slot, kind, value) in enumerate(fields):
        loc = tab_pos + offsets[i]
        if kind == "i8":
            struct.pack_into("<b", buf, loc, value)
        elif kind == "i32":
            struct.pack_into("<i", buf, loc, value)
        elif kind == "u32":
            struct.pack_into("<I", buf, loc, value)
        elif kind == "u64":
            struct.pack_into("<Q", buf, loc, value)
        elif kind == "str":
            patch(i, emit_string(value))
        elif kind in ("vec_u64", "vec_u8", "vec_struct16"):
            align = {"vec_u64": 8, "vec_u8": 1, "vec_struct16": 8}[kind]
            while (len(buf) + 4) % max(align, 4) or len(buf) % 4:
                buf.append(0)
            pos = len(buf)
            buf.extend(struct.pack("<I", len(value)))
            for v in value:
                if kind == "vec_u64":
                    buf.extend(struct.pack("<Q", v))
                elif kind == "vec_u8":
                    buf.extend(struct.pack("<B", v))
                else:
                    buf.extend(struct.pack("<IIQ", *v))
            patch(i, pos)
        elif kind == "vec_str":
            while len(buf) % 4:
                buf.append(0)
            pos = len(buf)
            buf.extend(struct.pack("<I", len(value)))
            buf.extend(bytes(4 * len(value)))
            for j, s in enumerate(value):
                spos = emit_string(s)
                eloc = pos + 4 + 4 * j
                struct.pack_into("<I", buf, eloc, spos - eloc)
            patch(i, pos)
        else:
            raise ValueError(kind)
    return bytes(buf)
