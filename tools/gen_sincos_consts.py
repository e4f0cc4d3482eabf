"""Print the double-double Taylor coefficients and the 3-part pi/2 split used by
aur_ppo_b200/csrc/det_sincos.h.  Run once; the output is pasted into the header."""
from mpmath import mp, mpf, factorial, pi
import struct
mp.prec = 400

def dd(v):
    hi = float(v)
    lo = float(v - mpf(hi))
    return hi, lo

def hx(d):
    return float.hex(d)

print("// sin: (sin(x)-x)/x^3 = S1 + S2 z + ...")
for k in range(1, 10):
    c = mpf((-1) ** k) / factorial(2 * k + 1)
    h, l = dd(c)
    print(f"S{k}: {hx(h)}, {hx(l)}")
print("// cos: (cos(x)-1+z/2)/z^2 = C2 + C3 z + ...")
for k in range(2, 11):
    c = mpf((-1) ** k) / factorial(2 * k)
    h, l = dd(c)
    print(f"C{k}: {hx(h)}, {hx(l)}")

# pi/2 in parts with 33 significant bits each (k * part exact for |k| < 2^20)
def trunc_bits(v, bits):
    m, e = mp.frexp(v)
    scaled = mp.floor(m * mpf(2) ** bits)
    return scaled * mpf(2) ** (e - bits)
rem = pi / 2
parts = []
for i in range(3):
    p = trunc_bits(rem, 33)
    parts.append(float(p)); assert mpf(float(p)) == p
    rem -= p
h, l = dd(rem)
print("PIO2_1..3:", [hx(p) for p in parts])
print("PIO2_tail dd:", hx(h), hx(l))
print("2/pi:", hx(float(2 / pi)))
print("pi/4:", hx(float(pi / 4)))
