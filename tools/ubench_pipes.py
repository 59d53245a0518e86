#!/usr/bin/env python3
"""Issue-rate microbenchmark for the sm_100a instruction mixes the histogram kernels are built
from.  Generates one kernel per mix (8 independent register chains, asm volatile), compiles it with
nvcc and -- on a GPU box -- prints warp-instructions per clock per SM sub-partition for 2/4/8 warps
per scheduler.  The question it answers: do ALU-pipe (LOP3/PRMT/SHF) and FMA-pipe instructions
overlap (bound = max of the pipes) or serialise on the dispatch port (bound = sum)?

    python tools/ubench_pipes.py build      # here: writes build/ubench/ubench_pipes(.cu), dumps SASS mnemonics
    python tools/ubench_pipes.py run        # on the GPU box: runs it, writes gpurun_out/ubench_pipes.txt
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "ubench")

# chain register classes: r = u32, f = f32, l = b64 (f32x2)
# each entry: name -> list of (class, ptx template); {d} = chain register, {a} {b} = loop-invariant operands of
# the same class held in registers
T = {
    "lop3": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;")],
    "prmt_reg": [("r", "prmt.b32 {d}, {a}, {b}, {d};")],
    "prmt_imm": [("r", "prmt.b32 {d}, {d}, {a}, 0x7440;")],
    "shf_funnel": [("r", "shf.r.wrap.b32 {d}, {d}, {a}, 4;")],
    "shl": [("r", "shl.b32 {d}, {d}, 4;")],
    "iadd": [("r", "add.s32 {d}, {d}, {a};")],
    "imad_reg": [("r", "mad.lo.s32 {d}, {d}, {a}, {b};")],
    "imad_imm": [("r", "mad.lo.s32 {d}, {d}, 4099, {a};")],
    "imad_pow2": [("r", "mad.lo.s32 {d}, {d}, 16, {a};")],
    "imad_wide": [("l", "mad.wide.u32 {d}, {a32}, {b32}, {d};")],
    "ffma_reg": [("f", "fma.rn.f32 {d}, {d}, {a}, {b};")],
    "ffma_imm": [("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};")],
    "ffma_imm_c": [("f", "fma.rn.f32 {d}, {a}, 0f3F800001, {d};")],
    "ffma_sat_imm": [("f", "fma.rn.sat.f32 {d}, {d}, 0f3F800001, {a};")],
    "ffma_rm_2imm": [("f", "fma.rm.f32 {d}, {d}, 0f3F800001, 0f4B000000;")],
    "fadd": [("f", "add.rn.f32 {d}, {d}, {a};")],
    "fmul_imm": [("f", "mul.rn.f32 {d}, {d}, 0f3F800001;")],
    "ffma2": [("l", "fma.rn.f32x2 {d}, {d}, {a}, {b};")],
    "fadd2": [("l", "add.rn.f32x2 {d}, {d}, {a};")],
    "hfma2": [("r", "fma.rn.f16x2 {d}, {d}, {a}, {b};")],
    "hfma2_sat": [("r", "fma.rn.sat.f16x2 {d}, {d}, {a}, {b};")],
    "fmnmx": [("f", "min.f32 {d}, {d}, {a};")],
    "imnmx": [("r", "min.s32 {d}, {d}, {a};")],
    "i2f_u32": [("r", "{{ .reg .f32 t; cvt.rn.f32.u32 t, {d}; mov.b32 {d}, t; }}")],
    "f2i_u32": [("r", "{{ .reg .f32 t; mov.b32 t, {d}; cvt.rzi.u32.f32 {d}, t; }}")],
    "popc": [("r", "popc.b32 {d}, {d};")],
    "mov": [("r", "mov.b32 {d}, {a};")],
    # mixes (one entry per instruction of the repeating pattern)
    "mix_lop3_ffma": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};")],
    "mix_lop3_2ffma": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};"),
                       ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};")],
    "mix_lop3_3ffma": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};"),
                       ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};"), ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};")],
    "mix_prmt_ffma": [("r", "prmt.b32 {d}, {a}, {b}, {d};"), ("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};")],
    "mix_lop3_ffmareg": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("f", "fma.rn.f32 {d}, {d}, {a}, {b};")],
    "mix_lop3_ffma2": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("l", "fma.rn.f32x2 {d}, {d}, {a}, {b};")],
    "mix_lop3_imad": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "mad.lo.s32 {d}, {d}, 4099, {a};")],
    "mix_ffma_imad": [("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};"), ("r", "mad.lo.s32 {d}, {d}, 4099, {a};")],
    "mix_ffma_ffma2": [("f", "fma.rn.f32 {d}, {d}, 0f3F800001, {a};"), ("l", "fma.rn.f32x2 {d}, {d}, {a}, {b};")],
    "mix_lop3_hfma2": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "fma.rn.f16x2 {d}, {d}, {a}, {b};")],
    "mix_lop3_prmt": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "prmt.b32 {d}, {a}, {b}, {d};")],
    "mix_lop3_f2i": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"),
                     ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"),
                     ("r", "{{ .reg .f32 t; mov.b32 t, {d}; cvt.rzi.u32.f32 {d}, t; }}")],
    "mix_lop3_popc": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"),
                      ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "popc.b32 {d}, {d};")],
    "mix_lop3_lds": [("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"),
                     ("r", "lop3.b32 {d}, {d}, {a}, {b}, 0x96;"), ("s", "")],
}
BODY = 96   # instructions per loop iteration (multiple of 1, 2, 3, 4)
CH = 8


def gen():
    src = ['#include <cstdio>', '#include <cstdint>', '#include <cuda_runtime.h>', '#include <vector>',
           '#include <string>', '#include <algorithm>', '']
    for name, pat in T.items():
        src.append(f'__global__ void __launch_bounds__(1024, 1) k_{name}(uint32_t* out, long long* cyc, int iters, '
                   'uint32_t ua, uint32_t ub) {')
        src.append('  __shared__ uint32_t sm[1024];')
        src.append('  sm[threadIdx.x] = threadIdx.x;')
        for c in range(CH):
            src.append(f'  uint32_t r{c} = threadIdx.x * {c + 3}u + ua; float f{c} = (float)(threadIdx.x + {c}) * 1e-3f; '
                       f'unsigned long long l{c} = ((unsigned long long)__float_as_uint(f{c}) << 32) | __float_as_uint(f{c});')
        src.append('  float fa = __uint_as_float(ua), fb = __uint_as_float(ub); uint32_t ra = ua, rb = ub;')
        src.append('  unsigned long long la = ((unsigned long long)ua << 32) | ua, lb = ((unsigned long long)ub << 32) | ub;')
        src.append('  __syncthreads();')
        src.append('  long long t0 = clock64();')
        src.append('  for (int it = 0; it < iters; ++it) {')
        for k in range(BODY):
            cls, tpl = pat[k % len(pat)]
            c = (k // len(pat)) % CH
            if cls == "s":
                src.append(f'    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r{c}) : "r"((uint32_t)__cvta_generic_to_shared(sm) + '
                           f'((threadIdx.x * 4u) & 4095u)), "r"(r{c}));')
                continue
            d = f"{cls}{c}"
            cons = {"r": "r", "f": "f", "l": "l"}[cls]
            txt = tpl.format(d="%0", a="%1", b="%2", a32="%3", b32="%4")
            if "{a32}" in tpl:
                src.append(f'    asm volatile("{txt}" : "+{cons}"({d}) : "{cons}"({cls}a), "{cons}"({cls}b), "r"(ua), "r"(ub));')
            else:
                src.append(f'    asm volatile("{txt}" : "+{cons}"({d}) : "{cons}"({cls}a), "{cons}"({cls}b));')
        src.append('  }')
        src.append('  long long t1 = clock64();')
        src.append('  __syncthreads();')
        src.append('  uint32_t acc = 0;')
        for c in range(CH):
            src.append(f'  acc ^= r{c} ^ __float_as_uint(f{c}) ^ (uint32_t)l{c} ^ (uint32_t)(l{c} >> 32);')
        src.append('  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;')
        src.append('  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;')
        src.append('}')
        src.append('')
    src.append('typedef void (*kfn)(uint32_t*, long long*, int, uint32_t, uint32_t);')
    src.append('struct Test { const char* name; kfn fn; };')
    src.append('static Test tests[] = {')
    for name in T:
        src.append(f'  {{"{name}", k_{name}}},')
    src.append('};')
    src.append(f'''
int main() {{
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, (size_t)nsm * 1024 * 4); cudaMalloc(&cyc, (size_t)nsm * 32 * 8);
  std::vector<long long> h(nsm * 32);
  const int iters = 2000;
  printf("%-20s %10s %10s %10s   (warp-instructions / clk / SMSP at 2, 4, 8 warps per scheduler)\\n", "mix", "w2", "w4", "w8");
  for (auto& t : tests) {{
    double ipc[3];
    for (int wi = 0; wi < 3; ++wi) {{
      int threads = 256 << wi;
      t.fn<<<nsm, threads>>>(out, cyc, 10, 0x3f800000u, 0x3f000000u);
      t.fn<<<nsm, threads>>>(out, cyc, iters, 0x3f800000u, 0x3f000000u);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {{ printf("%s: %s\\n", t.name, cudaGetErrorString(e)); return 1; }}
      cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost);
      int warps = threads / 32;
      std::vector<double> per;
      for (int b = 0; b < nsm; ++b) {{
        long long mx = 0;
        for (int w = 0; w < warps; ++w) mx = std::max(mx, h[b * 32 + w]);
        per.push_back((double)iters * {BODY} * warps / 4.0 / (double)mx);
      }}
      std::sort(per.begin(), per.end());
      ipc[wi] = per[per.size() / 2];
    }}
    printf("%-20s %10.3f %10.3f %10.3f\\n", t.name, ipc[0], ipc[1], ipc[2]);
  }}
  return 0;
}}
''')
    return "\n".join(src)


def build():
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(OUT, "ubench_pipes.cu")
    open(cu, "w").write(gen())
    exe = os.path.join(OUT, "ubench_pipes")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, cu])
    sass = subprocess.run(["cuobjdump", "-sass", exe], stdout=subprocess.PIPE, text=True).stdout
    # mnemonic histogram of every kernel's loop, to see what ptxas made of each PTX instruction
    cur, counts = None, {}
    for line in sass.splitlines():
        line = line.strip()
        if line.startswith("Function :"):
            cur = line.split(":")[1].strip()
            counts[cur] = {}
        elif cur and line.startswith("/*") and ";" in line:
            body = line.split("*/", 1)[1].strip()
            if body.startswith("@"):
                body = body.split(" ", 1)[1]
            mn = body.split(" ")[0].rstrip(";")
            counts[cur][mn] = counts[cur].get(mn, 0) + 1
    with open(os.path.join(OUT, "sass_mnemonics.txt"), "w") as f:
        for k, v in counts.items():
            top = sorted(v.items(), key=lambda kv: -kv[1])[:6]
            f.write(f"{k}: {top}\n")
    print(open(os.path.join(OUT, "sass_mnemonics.txt")).read())


def run():
    exe = os.path.join(OUT, "ubench_pipes")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(ROOT, "gpurun_out", "ubench_pipes.txt"), "w").write(r.stdout)
    print(r.stdout)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
