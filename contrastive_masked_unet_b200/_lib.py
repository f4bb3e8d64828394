"""ctypes binding of libcmu_b200.so (the C-ABI CUDA library, include/cmu_b200.h).

The prototypes are read from the header itself so that the Python side cannot drift from the C ABI.  There is NO
fallback: if the library is missing, cannot be loaded, or the device is not sm_100, every op raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcmu_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'cmu_b200.h')

_SCALARS = {
    'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float, 'double': ctypes.c_double,
    'unsigned int': ctypes.c_uint, 'unsigned long long': ctypes.c_ulonglong,
}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes], [argnames])} for every `cmu_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'^\s*#.*$', '', src, flags=re.M)          # preprocessor lines
    src = src.replace('extern "C" {', '')
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(cmu_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = ' '.join(m.group(1).split()), m.group(2), m.group(3).strip()
        if ret.replace(' ', '') == 'constchar*':
            restype = ctypes.c_char_p
        else:
            restype = _SCALARS[ret]
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                mm = re.match(r'(.*?)(\w+)$', a)
                typ, an = mm.group(1).strip(), mm.group(2)
                if '*' in typ:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_SCALARS[typ.replace('const ', '')])
                argnames.append(an)
        protos[name] = (restype, argtypes, argnames)
    return protos


# int-returning functions whose result is a value, not a status code
VALUE_FUNCS = {'cmu_version', 'cmu_mask_state_words', 'cmu_mask_workspace_bytes', 'cmu_conv3x3_c1_grid', 'cmu_conv_max_grid', 'cmu_bn_bwd_grid', 'cmu_bn_relu_head_grid'}


class CmuError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise CmuError(f'{LIB_PATH} is missing: build it with `python -m contrastive_masked_unet_b200.build` '
                           '(there is no CPU / PyTorch fallback for this path)')
        dll = ctypes.CDLL(LIB_PATH)
        self._protos = parse_header()
        for name, (restype, argtypes, _) in self._protos.items():
            fn = getattr(dll, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    @property
    def protos(self):
        if self._protos is None:
            self._protos = parse_header()
        return self._protos

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        dll = self.load()
        fn = getattr(dll, name)
        restype = self._protos[name][0]
        if restype is not ctypes.c_int or name in VALUE_FUNCS:
            return fn

        def checked(*args):
            rc = fn(*args)
            if rc != 0:
                raise CmuError(f'{name}: {dll.cmu_last_error().decode()}')
            return rc
        checked.__name__ = name
        self.__dict__[name] = checked
        return checked


lib = _Lib()
