"""commentjson.load / loads: JSON with '//' , '#' line comments and '/* */' blocks (outside strings) removed."""
import json


def _strip(text):
    out, i, n, in_str = [], 0, len(text), False
    while i < n:
        c = text[i]
        if in_str:
            out.append(c)
            if c == '\\' and i + 1 < n:
                out.append(text[i + 1])
                i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
            out.append(c)
        elif c == '#' or text.startswith('//', i):
            while i < n and text[i] != '\n':
                i += 1
            continue
        elif text.startswith('/*', i):
            j = text.find('*/', i + 2)
            i = n if j < 0 else j + 2
            continue
        else:
            out.append(c)
        i += 1
    return ''.join(out)


def loads(text, **kw):
    return json.loads(_strip(text), **kw)


def load(fp, **kw):
    return loads(fp.read(), **kw)


dumps, dump = json.dumps, json.dump
