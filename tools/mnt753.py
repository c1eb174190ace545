"""Pure-Python model of the MNT4753 / MNT6753 fields, towers and curves (arbitrary-precision ints).

Used by tools/gen_constants.py to derive every constant the CUDA/C++ code embeds, and by tests/ as a second,
independent checker next to oracle/ (C restatement) for the field / tower / curve formulas.
Numbers are the published curve parameters; the reference lists them at
depends/libff/libff/algebra/curves/mnt753/mnt4753/mnt4753_init.cpp:48-142,197-203 and
depends/libff/libff/algebra/curves/mnt753/mnt6753/mnt6753_init.cpp:50-155,213-219.
"""
LIMBS32 = 24
RBITS = 768
R = 1 << RBITS

# modulus A = Fr(MNT4753) = Fq(MNT6753); modulus B = Fq(MNT4753) = Fr(MNT6753)   
MOD_A = 0x1c4c62d92c41110229022eee2cdadb7f997505b8fafed5eb7e8f96c97d87307fdb925e8a0ed8d99d124d9a15af79db26c5c28c859a99b3eebca9429212636b9dff97634993aa4d6c381bc3f0057974ea099170fa13a4fd90776e240000001
MOD_B = 0x1c4c62d92c41110229022eee2cdadb7f997505b8fafed5eb7e8f96c97d87307fdb925e8a0ed8d99d124d9a15af79db117e776f218059db80f0da5cb537e38685acce9767254a4638810719ac425f0e39d54522cdd119f5e9063de245e8001

PRIMES = {"A": MOD_A, "B": MOD_B}
TWO_ADICITY = {"A": 30, "B": 15}
GENERATOR = 17  # multiplicative generator of both fields (mnt4753_init.cpp:68,94)


def root_of_unity(tag):
    """2^s-th primitive root = 17^((p-1)/2^s)  (equals libff's root_of_unity)."""
    p = PRIMES[tag]
    s = TWO_ADICITY[tag]
    return pow(GENERATOR, (p - 1) >> s, p)


class Curve:
    pass


def _mk():
    c4 = Curve()
    c4.name = "MNT4753"
    c4.fq_tag, c4.fr_tag = "B", "A"
    c4.q, c4.r = MOD_B, MOD_A
    c4.a = 2
    c4.b = 28798803903456388891410036793299405764940372360099938340752576406393880372126970068421383312482853541572780087363938442377933706865252053507077543420534380486492786626556269083255657125025963825610840222568694137138741554679540
    c4.ext_deg = 2
    c4.non_residue = 13
    c4.g1 = (23803503838482697364219212396100314255266282256287758532210460958670711284501374254909249084643549104668878996224193897061976788052185662569738774028756446662400954817676947337090686257134874703224133183061214213216866019444443,
             21091012152938225813050540665280291929032924333518476279110711148670464794818544820522390295209715531901248676888544060590943737249563733104806697968779796610374994498702698840169538725164956072726942500665132927942037078135054)
    c4.g2 = ((22367666623321080720060256844679369841450849258634485122226826668687008928557241162389052587294939105987791589807198701072089850184203060629036090027206884547397819080026926412256978135536735656049173059573120822105654153939204,
              19674349354065582663569886390557105215375764356464013910804136534831880915742161945711267871023918136941472003751075703860943205026648847064247080124670799190998395234694182621794580160576822167228187443851233972049521455293042),
             (6945425020677398967988875731588951175743495235863391886533295045397037605326535330657361771765903175481062759367498970743022872494546449436815843306838794729313050998681159000579427733029709987073254733976366326071957733646574,
              17406100775489352738678485154027036191618283163679980195193677896785273172506466216232026037788788436442188057889820014276378772936042638717710384987239430912364681046070625200474931975266875995282055499803236813013874788622488))
    # twist: a' = (a*13, 0), b' = (0, b*13)  (mnt4753_init.cpp:122-123)
    c4.twist_a = (c4.a * 13 % c4.q, 0)
    c4.twist_b = (0, c4.b * 13 % c4.q)

    c6 = Curve()
    c6.name = "MNT6753"
    c6.fq_tag, c6.fr_tag = "A", "B"
    c6.q, c6.r = MOD_A, MOD_B
    c6.a = 11
    c6.b = 11625908999541321152027340224010374716841167701783584648338908235410859267060079819722747939267925389062611062156601938166010098747920378738927832658133625454260115409075816187555055859490253375704728027944315501122723426879114
    c6.ext_deg = 3
    c6.non_residue = 11
    c6.g1 = (16364236387491689444759057944334173579070747473738339749093487337644739228935268157504218078126401066954815152892688541654726829424326599038522503517302466226143788988217410842672857564665527806044250003808514184274233938437290,
             4510127914410645922431074687553594593336087066778984214797709122300210966076979927285161950203037801392624582544098750667549188549761032654706830225743998064330900301346566408501390638273322467173741629353517809979540986561128)
    c6.g2 = ((46538297238006280434045879335349383221210789488441126073640895239023832290080310125413049878152095926176013036314720850781686614265244307536450228450615346834324267478485994670716807428718518299710702671895190475661871557310,
              10329739935427016564561842963551883445915701424214177782911128765230271790215029185795830999583638744119368571742929964793955375930677178544873424392910884024986348059137449389533744851691082159233065444766899262771358355816328,
              19962817058174334691864015232062671736353756221485896034072814261894530786568591431279230352444205682361463997175937973249929732063490256813101714586199642571344378012210374327764059557816647980334733538226843692316285591005879),
             (5648166377754359996653513138027891970842739892107427747585228022871109585680076240624013411622970109911154113378703562803827053335040877618934773712021441101121297691389632155906182656254145368668854360318258860716497525179898,
              26817850356025045630477313828875808893994935265863280918207940412617168254772789578700316551065949899971937475487458539503514034928974530432009759562975983077355912050606509147904958229398389093697494174311832813615564256810453,
              32332319709358578441696731586704495581796858962594701633932927358040566210788542624963749336109940335257143899293177116050031684054348958813290781394131284657165540476824211295508498842102093219808642563477603392470909217611033))
    # twist: a' = (0, 0, a), b' = (b*11, 0, 0)  (mnt6753_init.cpp:133-136)
    c6.twist_a = (0, 0, c6.a)
    c6.twist_b = (c6.b * 11 % c6.q, 0, 0)
    return c4, c6


MNT4753, MNT6753 = _mk()
CURVES = {"MNT4753": MNT4753, "MNT6753": MNT6753}


# ---------------------------------------------------------------- Montgomery encoding helpers
def to_mont(x, p):
    return (x << RBITS) % p


def from_mont(x, p):
    return x * pow(R, -1, p) % p


def to_limbs32(x, n=LIMBS32):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_bytes(b):
    return int.from_bytes(b, "little")


def to_bytes(x, n=96):
    return int(x).to_bytes(n, "little")


# ---------------------------------------------------------------- extension-field arithmetic (tuples of ints)
class ExtField:
    """Fq[u]/(u^deg - nr), elements are tuples. deg=1 degenerates to the base field."""

    def __init__(self, p, deg, nr):
        self.p, self.deg, self.nr = p, deg, nr

    def zero(self):
        return (0,) * self.deg

    def one(self):
        return (1,) + (0,) * (self.deg - 1)

    def add(self, a, b):
        return tuple((x + y) % self.p for x, y in zip(a, b))

    def sub(self, a, b):
        return tuple((x - y) % self.p for x, y in zip(a, b))

    def neg(self, a):
        return tuple((-x) % self.p for x in a)

    def mul(self, a, b):
        d = self.deg
        t = [0] * (2 * d - 1)
        for i in range(d):
            for j in range(d):
                t[i + j] += a[i] * b[j]
        for k in range(2 * d - 2, d - 1, -1):
            t[k - d] += self.nr * t[k]
        return tuple(x % self.p for x in t[:d])

    def sqr(self, a):
        return self.mul(a, a)

    def is_zero(self, a):
        return all(x == 0 for x in a)

    def inv(self, a):
        p, nr = self.p, self.nr
        if self.deg == 1:
            return (pow(a[0], -1, p),)
        if self.deg == 2:  # 1/(a+bu) = (a-bu)/(a^2 - nr b^2)
            t = pow((a[0] * a[0] - nr * a[1] * a[1]) % p, -1, p)
            return (a[0] * t % p, (-a[1] * t) % p)
        x, y, z = a  # deg 3: adjugate / norm
        c0 = (x * x - nr * y * z) % p
        c1 = (nr * z * z - x * y) % p
        c2 = (y * y - x * z) % p
        t = pow((x * c0 + nr * (z * c1 + y * c2)) % p, -1, p)
        return (c0 * t % p, c1 * t % p, c2 * t % p)


def g1_field(c):
    return ExtField(c.q, 1, 0)


def g2_field(c):
    return ExtField(c.q, c.ext_deg, c.non_residue)


# ---------------------------------------------------------------- affine curve arithmetic (None = infinity)
def ec_add(F, a_coeff, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if y1 != y2 or F.is_zero(y1):
            return None
        num = F.add(F.add(F.add(F.sqr(x1), F.sqr(x1)), F.sqr(x1)), a_coeff)
        den = F.add(y1, y1)
    else:
        num = F.sub(y2, y1)
        den = F.sub(x2, x1)
    lam = F.mul(num, F.inv(den))
    x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_neg(F, P):
    return None if P is None else (P[0], F.neg(P[1]))


def ec_mul(F, a_coeff, k, P):
    R_ = None
    for bit in bin(k)[2:] if k else "":
        R_ = ec_add(F, a_coeff, R_, R_)
        if bit == "1":
            R_ = ec_add(F, a_coeff, R_, P)
    return R_


def on_curve(F, a_coeff, b_coeff, P):
    if P is None:
        return True
    x, y = P
    return F.sqr(y) == F.add(F.add(F.mul(F.sqr(x), x), F.mul(a_coeff, x)), b_coeff)


def g1_params(c):
    F = g1_field(c)
    return F, (c.a,), (c.b,), ((c.g1[0],), (c.g1[1],))


def g2_params(c):
    F = g2_field(c)
    return F, tuple(c.twist_a), tuple(c.twist_b), (tuple(c.g2[0]), tuple(c.g2[1]))


if __name__ == "__main__":
    for c in (MNT4753, MNT6753):
        F, a, b, G = g1_params(c)
        assert on_curve(F, a, b, G), c.name
        assert ec_mul(F, a, c.r, G) is None
        F, a, b, G = g2_params(c)
        assert on_curve(F, a, b, G), c.name + " g2"
        assert ec_mul(F, a, c.r, G) is None
    for t in "AB":
        w = root_of_unity(t)
        assert pow(w, 1 << TWO_ADICITY[t], PRIMES[t]) == 1 and pow(w, 1 << (TWO_ADICITY[t] - 1), PRIMES[t]) != 1
    print("mnt753.py self-check ok")
