"""NRRD container I/O for the episode builder ("next" row N4 of SURVEY §8f).

The reference reads its volumes with the third-party `pynrrd` package (`nrrd.read(path)` in
dataset/few_shot_reader.py:327,335 and dataset/brain_reader.py; pynrrd is not vendored in the reference, carries no version
pin there and is absent from this image), so this module restates the published NRRD format (NRRD0001..0005: a text header
of `field: value` lines and `key:=value` pairs, a blank line, then the payload) with pynrrd's calling convention:

    data, header = read(path)          # index_order='F': data.shape == header['sizes'] (fastest axis first), like pynrrd
    write(path, data, header=None)     # NRRD0004, little endian, gzip (or raw) payload

Supported: every scalar `type` alias of the format, `endian`, encodings raw / gzip (gz) / bzip2 (bz2) / ascii (text, txt),
`line skip` / `byte skip` (including byte skip -1 for raw), attached and detached (`data file:`) payloads.  Not supported:
`block` type and multi-file detached lists (`data file: LIST` / printf patterns) — they raise NrrdError.
"parity unpinned" for the container itself (no pynrrd here to compare with); tests/test_dataset.py checks the reader against
hand-assembled byte strings and round trips, and the readers built on it against the reference's Dataset classes."""
import bz2
import os
import zlib
from collections import OrderedDict

import numpy as np


class NrrdError(ValueError):
    pass


_TYPES = {}
for _names, _code in (
        (('signed char', 'int8', 'int8_t'), 'i1'),
        (('uchar', 'unsigned char', 'uint8', 'uint8_t'), 'u1'),
        (('short', 'short int', 'signed short', 'signed short int', 'int16', 'int16_t'), 'i2'),
        (('ushort', 'unsigned short', 'unsigned short int', 'uint16', 'uint16_t'), 'u2'),
        (('int', 'signed int', 'int32', 'int32_t'), 'i4'),
        (('uint', 'unsigned int', 'uint32', 'uint32_t'), 'u4'),
        (('longlong', 'long long', 'long long int', 'signed long long', 'signed long long int', 'int64', 'int64_t'), 'i8'),
        (('ulonglong', 'unsigned long long', 'unsigned long long int', 'uint64', 'uint64_t'), 'u8'),
        (('float',), 'f4'),
        (('double',), 'f8')):
    for _n in _names:
        _TYPES[_n] = _code
_TYPE_NAME = {'i1': 'int8', 'u1': 'uint8', 'i2': 'int16', 'u2': 'uint16', 'i4': 'int32', 'u4': 'uint32', 'i8': 'int64',
              'u8': 'uint64', 'f4': 'float', 'f8': 'double'}
_ALIASES = {'datafile': 'data file', 'lineskip': 'line skip', 'byteskip': 'byte skip', 'centers': 'centerings',
            'axismins': 'axis mins', 'axismaxs': 'axis maxs', 'oldmin': 'old min', 'oldmax': 'old max',
            'blocksize': 'block size', 'sampleunits': 'sample units'}
_INT_FIELDS = {'dimension', 'line skip', 'byte skip', 'space dimension', 'block size'}
_INT_VECTOR_FIELDS = {'sizes'}
_FLOAT_VECTOR_FIELDS = {'spacings', 'thicknesses', 'axis mins', 'axis maxs'}


def _parse_vector(text):
    text = text.strip()
    if text == 'none':
        return None
    if not (text.startswith('(') and text.endswith(')')):
        raise NrrdError('malformed NRRD vector %r' % text)
    return np.array([float(v) for v in text[1:-1].split(',')], dtype=np.float64)


def _parse_field(name, value):
    if name in _INT_FIELDS:
        return int(value)
    if name in _INT_VECTOR_FIELDS:
        return np.array([int(v) for v in value.split()], dtype=np.int64)
    if name in _FLOAT_VECTOR_FIELDS:
        return np.array([float(v) for v in value.split()], dtype=np.float64)
    if name == 'space origin':
        return _parse_vector(value)
    if name == 'space directions':
        rows = [_parse_vector(v) for v in value.split()]
        width = max((len(r) for r in rows if r is not None), default=0)
        return np.array([r if r is not None else np.full(width, np.nan) for r in rows], dtype=np.float64)
    if name in ('kinds', 'centerings'):
        return value.split()
    if name in ('type', 'encoding', 'endian'):
        return value.lower()
    return value


def read_header(stream):
    """Parses the header from a binary stream positioned at byte 0; leaves the stream at the first payload byte."""
    magic = stream.readline().decode('ascii', 'ignore').rstrip('\r\n')
    if not (magic.startswith('NRRD000') and magic[7:].isdigit() and 1 <= int(magic[7:]) <= 5):
        raise NrrdError('not an NRRD file (magic line %r)' % magic)
    header = OrderedDict()
    while True:
        raw = stream.readline()
        if raw == b'':                                  # header-only file (detached payload) may end without a blank line
            break
        line = raw.decode('ascii', 'ignore').rstrip('\r\n')
        if line == '':
            break
        if line.startswith('#'):
            continue
        if ':=' in line and (': ' not in line or line.index(':=') < line.index(': ')):
            key, value = line.split(':=', 1)
            header[key] = value.replace('\\n', '\n').replace('\\\\', '\\')
            continue
        if ': ' not in line and not line.endswith(':'):
            raise NrrdError('malformed NRRD header line %r' % line)
        name, _, value = line.partition(':')
        name = name.strip().lower()
        name = _ALIASES.get(name, name)
        if name in header:
            raise NrrdError('duplicate NRRD header field %r' % name)
        header[name] = _parse_field(name, value.strip())
    for need in ('type', 'dimension', 'sizes', 'encoding'):
        if need not in header:
            raise NrrdError('NRRD header lacks the required field %r' % need)
    if len(header['sizes']) != header['dimension']:
        raise NrrdError('NRRD sizes %s do not match dimension %d' % (header['sizes'], header['dimension']))
    return header


def _dtype_of(header):
    kind = header['type']
    if kind == 'block':
        raise NrrdError("NRRD type 'block' is not supported")
    if kind not in _TYPES:
        raise NrrdError('unknown NRRD type %r' % kind)
    code = _TYPES[kind]
    if code[1] == '1':
        return np.dtype(code)
    endian = header.get('endian')
    if endian not in ('little', 'big'):
        raise NrrdError("NRRD header needs 'endian: little|big' for multi-byte type %r" % kind)
    return np.dtype(('<' if endian == 'little' else '>') + code)


def read(path):
    """pynrrd-compatible read: returns (data, header) with data.shape == tuple(header['sizes'])."""
    with open(path, 'rb') as fh:
        header = read_header(fh)
        dtype = _dtype_of(header)
        count = int(np.prod(header['sizes']))
        encoding = header['encoding']
        line_skip, byte_skip = header.get('line skip', 0), header.get('byte skip', 0)
        if 'data file' in header:
            target = header['data file']
            if target.split()[0] == 'LIST' or '%' in target:
                raise NrrdError('multi-file detached NRRD payloads are not supported')
            if not os.path.isabs(target):
                target = os.path.join(os.path.dirname(os.path.abspath(path)), target)
            with open(target, 'rb') as dfh:
                payload = dfh.read()
        else:
            payload = fh.read()
    # `line skip` / `byte skip` act on the stored bytes for raw, and — as in the format text — on the decoded stream for the
    # compressed encodings; byte skip -1 means "the payload is the tail of the file" and is defined for raw only
    def skip(buf):
        pos = 0
        for _ in range(line_skip):
            nl = buf.find(b'\n', pos)
            if nl < 0:
                raise NrrdError('line skip runs past the end of the NRRD payload')
            pos = nl + 1
        if byte_skip > 0:
            pos += byte_skip
        return buf[pos:]
    if encoding == 'raw':
        if byte_skip == -1:
            payload = payload[len(payload) - count * dtype.itemsize:]
        else:
            payload = skip(payload)
        if len(payload) < count * dtype.itemsize:
            raise NrrdError('NRRD payload has %d bytes, %d expected' % (len(payload), count * dtype.itemsize))
        flat = np.frombuffer(payload, dtype=dtype, count=count)
    elif encoding in ('gzip', 'gz', 'bzip2', 'bz2'):
        if byte_skip == -1:
            raise NrrdError('byte skip -1 is only valid with raw encoding')
        try:
            raw = zlib.decompress(payload, 32 + zlib.MAX_WBITS) if encoding in ('gzip', 'gz') else bz2.decompress(payload)
        except (zlib.error, OSError, ValueError) as e:
            raise NrrdError('cannot decompress the NRRD payload: %s' % e)
        raw = skip(raw)
        if len(raw) < count * dtype.itemsize:
            raise NrrdError('NRRD payload has %d bytes, %d expected' % (len(raw), count * dtype.itemsize))
        flat = np.frombuffer(raw, dtype=dtype, count=count)
    elif encoding in ('ascii', 'text', 'txt'):
        values = skip(payload).split()
        if len(values) < count:
            raise NrrdError('NRRD text payload has %d values, %d expected' % (len(values), count))
        flat = np.array([float(v) for v in values[:count]]).astype(dtype)
    else:
        raise NrrdError('unsupported NRRD encoding %r' % encoding)
    # the first axis of `sizes` is the fastest one in the file
    data = flat.reshape(tuple(int(s) for s in header['sizes']), order='F')
    if not data.flags.writeable:                         # frombuffer views are read-only; pynrrd hands back a writable array
        data = data.copy(order='F')
    return data, header


def _format_value(name, value):
    if isinstance(value, np.ndarray) and value.ndim == 2:
        return ' '.join('none' if np.all(np.isnan(r)) else '(' + ','.join(repr(float(v)) for v in r) + ')' for r in value)
    if name == 'space origin':
        return '(' + ','.join(repr(float(v)) for v in value) + ')'
    if isinstance(value, (list, tuple, np.ndarray)):
        return ' '.join(str(v) for v in value)
    return str(value)


def write(path, data, header=None, encoding='gzip'):
    """Writes `data` (any supported scalar dtype) with data.shape as `sizes`; extra `header` fields are carried through."""
    data = np.asarray(data)
    code = data.dtype.kind + str(data.dtype.itemsize)
    if code not in _TYPE_NAME:
        raise NrrdError('dtype %s has no NRRD type' % data.dtype)
    fields = OrderedDict()
    fields['type'] = _TYPE_NAME[code]
    fields['dimension'] = data.ndim
    fields['sizes'] = list(data.shape)
    if data.dtype.itemsize > 1:
        fields['endian'] = 'little'
    fields['encoding'] = encoding
    pairs = OrderedDict()
    for k, v in (header or {}).items():
        if k in ('type', 'dimension', 'sizes', 'endian', 'encoding', 'data file', 'line skip', 'byte skip'):
            continue
        (fields if k == k.lower() and ':=' not in k and k in _KNOWN_FIELDS else pairs)[k] = v
    payload = np.asfortranarray(data.astype(data.dtype.newbyteorder('<'), copy=False)).tobytes(order='F')
    if encoding in ('gzip', 'gz'):
        comp = zlib.compressobj(6, zlib.DEFLATED, 16 + zlib.MAX_WBITS)
        payload = comp.compress(payload) + comp.flush()
    elif encoding in ('bzip2', 'bz2'):
        payload = bz2.compress(payload)
    elif encoding != 'raw':
        raise NrrdError('write supports raw, gzip and bzip2 encodings (got %r)' % encoding)
    with open(path, 'wb') as fh:
        fh.write(b'NRRD0004\n# written by rpnet_b200.dataset.nrrd_io\n')
        for k, v in fields.items():
            fh.write(('%s: %s\n' % (k, _format_value(k, v))).encode('ascii'))
        for k, v in pairs.items():
            fh.write(('%s:=%s\n' % (k, str(v).replace('\\', '\\\\').replace('\n', '\\n'))).encode('ascii'))
        fh.write(b'\n')
        fh.write(payload)


_KNOWN_FIELDS = {'space', 'space dimension', 'space units', 'space origin', 'space directions', 'measurement frame', 'kinds',
                 'centerings', 'spacings', 'thicknesses', 'axis mins', 'axis maxs', 'labels', 'units', 'min', 'max', 'old min',
                 'old max', 'content', 'sample units', 'number'}
