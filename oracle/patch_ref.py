#!/usr/bin/env python
"""TEST INFRASTRUCTURE: writes PATCHED COPIES of two reference files into oracle/_ref/patched/ (git-ignored) -- the three
edits INTEGRATION.md section 1 asks a maintainer to make so that "CUDA" is a renderer type name and Ray::CreateRenderer
tries the CUDA backend first.  The reference tree itself is never modified; nothing is copied into the repository."""
import os
import sys

ref, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)


def patch(name, edits):
    s = open(os.path.join(ref, name)).read()
    for anchor, new in edits:
        if s.count(anchor) != 1:
            sys.exit(f"patch_ref: anchor not found exactly once in {name}: {anchor!r}")
        s = s.replace(anchor, new)
    open(os.path.join(out, name), "w").write(s)


patch("RendererBase.cpp", [
    ('    case eRendererType::DirectX12:\n        return "DX";\n',
     '    case eRendererType::DirectX12:\n        return "DX";\n    case eRendererType(8):\n        return "CUDA";\n'),
    ('    } else if (name == "DX") {\n        return eRendererType::DirectX12;\n    }\n',
     '    } else if (name == "DX") {\n        return eRendererType::DirectX12;\n    } else if (name == "CUDA") {\n'
     '        return eRendererType(8);\n    }\n'),
])
patch("Ray.cpp", [
    ('#if defined(ENABLE_VK_IMPL)\n    if (enabled_types & eRendererType::Vulkan) {',
     '    if (enabled_types & eRendererType(8)) {\n'
     '        log->Info("Ray: Creating CUDA renderer %ix%i", s.w, s.h);\n'
     '        try {\n'
     '            return Cuda::CreateRenderer(s, log);\n'
     '        } catch (std::exception &e) {\n'
     '            log->Info("Ray: Failed to create CUDA renderer, %s", e.what());\n'
     '        }\n'
     '    }\n'
     '#if defined(ENABLE_VK_IMPL)\n    if (enabled_types & eRendererType::Vulkan) {'),
    ('namespace Ray {\nLogNull g_null_log;',
     'namespace Ray {\nnamespace Cuda {\nRendererBase *CreateRenderer(const settings_t &s, ILog *log);\n}\nLogNull g_null_log;'),
])
print("patched copies written to", out)
