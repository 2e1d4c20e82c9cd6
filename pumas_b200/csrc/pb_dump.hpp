/*
 * pb_dump.hpp -- host side loader of a `pumas_physics_dump` (host only, no CUDA).
 *
 * The dump format (src/pumas.c:2167-2196) is: int tag (13), int size, then
 * `size` raw bytes of `struct pumas_physics` (pumas.c:660-806, LP64 layout)
 * followed by its data segment. The 53 data pointers that start at `mdf_path`
 * hold stale absolute addresses; they are turned into offsets exactly like the
 * reference rebases them (pumas.c:2096-2122). Nothing is recomputed: tables are
 * used bit for bit.
 *
 * The loader builds one relocatable image (the raw blob followed by a few
 * derived, flattened arrays) so that the same `pb::Physics` view can be
 * instantiated over the host copy and over the copy in HBM.
 */
#ifndef PB_DUMP_HPP
#define PB_DUMP_HPP

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pb_atomic_data.h"
#include "pb_csda_tables.h"
#include "pb_hot_tables.h"
#include "pb_physics.h"

namespace pb {

constexpr int PHYSICS_BINARY_DUMP_TAG = 13; /* pumas.c:665 */

/* Mirror of the head of struct pumas_physics (pumas.c:667-805), LP64. */
struct DumpHeader {
        int32_t size;
        int32_t n_energies;
        int32_t n_materials;
        int32_t n_composites;
        int32_t n_elements;
        int32_t n_components;
        int32_t max_components;
        int32_t n_energy_loss_header;
        int32_t dcs_model_offset;
        int32_t n_table_dcs;
        int32_t particle;
        int32_t pad_;
        double ctau;
        double mass;
        double cutoff;
        double elastic_ratio;
        /* The N_DATA_POINTERS = 53 rebased pointers, in declaration order. */
        uint64_t mdf_path, dedx_path, dedx_filename;
        uint64_t table_K, table_K_dX, table_K_dNI_in, table_K_dNI_el;
        uint64_t table_X, table_X_dK, table_X_dT;
        uint64_t table_T, table_T_dK;
        uint64_t table_dE, table_dE_dK;
        uint64_t table_Omega, table_Omega_dK;
        uint64_t table_NI_el, table_NI_el_dK;
        uint64_t table_NI_in, table_NI_in_dK;
        uint64_t table_CS, table_CS_dK;
        uint64_t table_CSf, table_CSf_dK;
        uint64_t table_CSn, table_CSn_dK;
        uint64_t table_DCS, table_DCS_x, table_DCS_envelope;
        uint64_t table_Xt, table_Kt;
        uint64_t table_Li, table_Li_dK;
        uint64_t table_a_max, table_b_max;
        uint64_t table_Mu0, table_Mu0_dK;
        uint64_t table_Lb, table_Lb_dK;
        uint64_t table_Ms1, table_Ms1_dK;
        uint64_t elements_in;
        uint64_t material_density, material_ZoA, material_I, material_aS;
        uint64_t element, composition, composite, material_name;
        uint64_t model_bremsstrahlung, model_pair_production,
            model_photonuclear;
        /* Stale function pointers, ignored. */
        uint64_t dcs_bremsstrahlung, dcs_pair_production, dcs_photonuclear;
};
static_assert(sizeof(DumpHeader) == 528, "struct pumas_physics head layout");

/* struct atomic_element (pumas.c:510-523) */
struct DumpElement {
        double Z, A, I;
        int32_t index;
        int32_t pad_;
        uint64_t name;
};
static_assert(sizeof(DumpElement) == 40, "struct atomic_element layout");

/* struct material_component (pumas.c:527-532) */
struct DumpComponent {
        int32_t element;
        int32_t pad_;
        double fraction;
};
static_assert(sizeof(DumpComponent) == 16, "struct material_component layout");

struct HostPhysics {
        /* tag + size + blob, exactly as read (pumas_physics_dump writes it
         * back unchanged). */
        std::vector<unsigned char> raw;
        /* Relocatable image: blob, then the derived arrays. */
        std::vector<unsigned char> image;
        /* View over `image` (host pointers). */
        Physics view;
        /* Offsets of every pointer member of `view` inside `image`. */
        std::vector<std::pair<size_t, size_t> > relocs; /* (member, offset) */

        std::vector<std::string> material_name;
        std::vector<std::string> element_name;
        std::vector<double> material_I;
        /* struct composite_material of the composites (pumas.c:536-550):
         * base material index and mass fraction of every component. */
        std::vector<std::vector<std::pair<int, double> > > composite;
        std::string model[3];
        int model_index[3];
        int n_energy_loss_header;

        /* Instantiate the view over a copy of `image` located at `base`. */
        Physics view_at(const void * base) const
        {
                Physics v = view;
                unsigned char * pv = reinterpret_cast<unsigned char *>(&v);
                for (size_t i = 0; i < relocs.size(); i++) {
                        const unsigned char * addr =
                            static_cast<const unsigned char *>(base) +
                            relocs[i].second;
                        std::memcpy(pv + relocs[i].first, &addr, sizeof addr);
                }
                return v;
        }
};

namespace detail {

inline size_t align16(size_t n) { return (n + 15) & ~size_t(15); }

/* Fold the Z-indexed published tables into the per element constants. Same
 * selection logic as pumas.c:7143-7175, 7351-7370, 13709-13738. */
inline void element_fill(Element & e, double Z, double A, double I)
{
        std::memset(&e, 0, sizeof e);
        e.Z = Z;
        e.A = A;
        e.I = I;

        int iZ = (int)Z - 1;
        const int nZ = 103;
        if (iZ >= nZ) iZ = nZ - 1;
        if (iZ < 0) iZ = 0;
        const double * row = data::salvat_screening[iZ];
        if (row[4] == 1.) {
                if (iZ == 0) {
                        e.n_screening = 1;
                        e.amplitude[0] = 1.;
                } else {
                        e.n_screening = 2;
                        e.amplitude[0] = row[0];
                        e.amplitude[1] = 1. - e.amplitude[0];
                }
        } else {
                e.n_screening = 3;
                e.amplitude[0] = row[0];
                e.amplitude[1] = row[1];
                e.amplitude[2] = 1. - (e.amplitude[0] + e.amplitude[1]);
        }
        for (int i = 0; i < 3; i++) e.alpha_over_a0[i] = row[2 + i] / BOHR_RADIUS;

        int iN = (int)Z;
        const int nN = 120;
        if (iN >= nN - 1)
                iN = nN - 2;
        else if ((iN == 1) && (A >= 1.5))
                iN = 0;
        else if ((iN == 11) && (A == 22.))
                iN = nN - 1;
        if (iN < 0) iN = 0;
        e.nuclear_radius = data::nuclear_radius_fm[iN] * 1E-15;

        e.radiation_log = 182.7;
        {
                const int i = (int)Z;
                if ((i >= 1) && (i <= 92)) {
                        const double l = data::radiation_logarithm[i - 1];
                        if (l > 0) e.radiation_log = l;
                }
        }

        /* volatile: evaluate with the run time libm, like the reference. */
        volatile double third = 1. / 3., mthird = -1. / 3., twothird = 2. / 3.;
        volatile double e027 = 0.27;
        e.z13 = pow(Z, third);
        e.z13_inv = pow(Z, mthird);
        e.z23 = pow(Z, twothird);
        e.Dn = 1.54 * pow(A, e027);
        e.lnA = log(A);
        e.a13 = pow(A, third);
        if (Z == 1) {
                e.wood_saxon = 0.;
        } else {
                /* dcs_wood_saxon, pumas.c:14732-14748 */
                volatile double two = 2.;
                double r0 = e.a13;
                r0 = 1.12 * r0 - 0.86 / r0;
                const double a = 0.54;
                const double zeta3 = 1.202056903159594;
                const double V = a *
                    (r0 * r0 * log(two) + a * r0 * M_PI * M_PI / 6. +
                        a * a * 1.5 * zeta3);
                e.wood_saxon = 1. - 4. * M_PI * 0.17 * V / A;
        }
}

} /* namespace detail */

/* Parse a dump held in memory. Returns 0 or an enum pumas_return code, with a
 * message in `err`. */
inline int load_physics_bytes(
    const unsigned char * bytes, size_t n_bytes, HostPhysics & hp,
    std::string & err)
{
        if (n_bytes < 8) {
                err = "could not read the dump header";
                return 10; /* PUMAS_RETURN_IO_ERROR */
        }
        int32_t tag, size;
        std::memcpy(&tag, bytes, 4);
        std::memcpy(&size, bytes + 4, 4);
        if (tag != PHYSICS_BINARY_DUMP_TAG) {
                err = "incompatible version of binary dump";
                return 11; /* PUMAS_RETURN_FORMAT_ERROR */
        }
        if ((size < (int)sizeof(DumpHeader)) || (n_bytes < 8 + (size_t)size)) {
                err = "truncated binary dump";
                return 10;
        }
        hp.raw.assign(bytes, bytes + 8 + size);
        const unsigned char * blob = hp.raw.data() + 8;
        DumpHeader h;
        std::memcpy(&h, blob, sizeof h);
        if (h.size != size) {
                err = "inconsistent size in binary dump";
                return 11;
        }

        /* A stored pointer -> offset inside the blob (pumas.c:2096-2100). */
        const uint64_t base = h.mdf_path;
        const size_t data0 = sizeof(DumpHeader);
        bool ok = true;
        auto off = [&](uint64_t ptr, size_t need) -> size_t {
                const uint64_t o = ptr - base + data0;
                if ((ptr < base) || (o + need > (uint64_t)size)) {
                        ok = false;
                        return data0;
                }
                return (size_t)o;
        };

        const size_t ne = h.n_energies, nm = h.n_materials,
                     nel = h.n_elements, nc = h.n_components;
        if ((h.n_energies < 3) || (h.n_materials < 1) || (h.n_elements < 1) ||
            (h.n_components < 1) || (h.n_table_dcs < 2)) {
                err = "corrupted binary dump (bad dimensions)";
                return 11;
        }

        /* Model names -> indices (pb_dcs_models.h; dcs_stack,
         * pumas.c:15381-15390). */
        const uint64_t model_ptr[3] = { h.model_bremsstrahlung,
                h.model_pair_production, h.model_photonuclear };
        const char * known[3][3] = { { "SSR", "KKP", "ABB" },
                { "SSR", "KKP", nullptr }, { "DRSS", "BM", "BBKS" } };
        for (int i = 0; i < 3; i++) {
                const size_t o = off(model_ptr[i], 1);
                if (!ok) break;
                hp.model[i] = std::string(
                    reinterpret_cast<const char *>(blob + o),
                    strnlen(reinterpret_cast<const char *>(blob + o),
                        size - o));
                hp.model_index[i] = -1;
                for (int k = 0; k < 3; k++)
                        if ((known[i][k] != nullptr) &&
                            (hp.model[i] == known[i][k]))
                                hp.model_index[i] = k;
                /* (anything else: a user registered model, resolved by the
                 * caller against its registry; pumas.c:15422) */
                if (hp.model_index[i] < 0) hp.model_index[i] = MODEL_USER;
        }

        /* Elements. */
        std::vector<Element> elements(nel);
        hp.element_name.resize(nel);
        {
                const size_t otab = off(h.element, nel * 8);
                for (size_t i = 0; ok && (i < nel); i++) {
                        uint64_t pe;
                        std::memcpy(&pe, blob + otab + 8 * i, 8);
                        const size_t oe = off(pe, sizeof(DumpElement));
                        if (!ok) break;
                        DumpElement de;
                        std::memcpy(&de, blob + oe, sizeof de);
                        if (de.index != (int)i) {
                                err = "unexpected element index in dump";
                                return 11;
                        }
                        detail::element_fill(elements[i], de.Z, de.A, de.I);
                        const size_t on = off(de.name, 1);
                        if (!ok) break;
                        hp.element_name[i] = std::string(
                            reinterpret_cast<const char *>(blob + on),
                            strnlen(reinterpret_cast<const char *>(blob + on),
                                size - on));
                }
        }

        /* Compositions, flattened in material order. */
        std::vector<int32_t> elements_in(nm);
        if (ok)
                std::memcpy(elements_in.data(),
                    blob + off(h.elements_in, nm * 4), nm * 4);
        std::vector<int32_t> comp_offset(nm + 1, 0), comp_element;
        std::vector<double> comp_fraction;
        {
                const size_t otab = off(h.composition, nm * 8);
                for (size_t m = 0; ok && (m < nm); m++) {
                        const int n_in = elements_in[m];
                        if ((n_in < 1) || (n_in > h.max_components)) {
                                err = "corrupted binary dump (components)";
                                return 11;
                        }
                        uint64_t pc;
                        std::memcpy(&pc, blob + otab + 8 * m, 8);
                        const size_t oc = off(pc, n_in * sizeof(DumpComponent));
                        if (!ok) break;
                        for (int i = 0; i < n_in; i++) {
                                DumpComponent dc;
                                std::memcpy(&dc,
                                    blob + oc + i * sizeof(DumpComponent),
                                    sizeof dc);
                                if ((dc.element < 0) ||
                                    (dc.element >= (int)nel)) {
                                        err = "corrupted binary dump "
                                              "(element index)";
                                        return 11;
                                }
                                comp_element.push_back(dc.element);
                                comp_fraction.push_back(dc.fraction);
                        }
                        comp_offset[m + 1] = (int32_t)comp_element.size();
                }
                if (ok && (comp_element.size() != nc)) {
                        err = "corrupted binary dump (n_components)";
                        return 11;
                }
        }

        /* Composite materials: `composite` is an array of pointers to
         * { int n_components; struct { int material; double fraction; }[] }
         * (pumas.c:536-550, rebased like pumas.c:2115-2119). Host side
         * information only (pumas_physics_composite_properties). */
        hp.composite.clear();
        if (h.n_composites > 0) {
                const size_t ncomp = (size_t)h.n_composites;
                if (ncomp > nm) {
                        err = "corrupted binary dump (n_composites)";
                        return 11;
                }
                const size_t otab = off(h.composite, ncomp * 8);
                for (size_t k = 0; ok && (k < ncomp); k++) {
                        uint64_t pc;
                        std::memcpy(&pc, blob + otab + 8 * k, 8);
                        const size_t o = off(pc, 8);
                        if (!ok) break;
                        int32_t n_in;
                        std::memcpy(&n_in, blob + o, 4);
                        if ((n_in < 1) || ((size_t)n_in > nm)) {
                                err = "corrupted binary dump (composite)";
                                return 11;
                        }
                        const size_t oc = off(pc, 8 + (size_t)n_in * 16);
                        if (!ok) break;
                        std::vector<std::pair<int, double> > parts;
                        for (int i = 0; i < n_in; i++) {
                                int32_t material;
                                double fraction;
                                std::memcpy(&material,
                                    blob + oc + 8 + 16 * i, 4);
                                std::memcpy(&fraction,
                                    blob + oc + 16 + 16 * i, 8);
                                parts.push_back(
                                    std::make_pair((int)material, fraction));
                        }
                        hp.composite.push_back(parts);
                }
        }

        /* Names. */
        hp.material_name.resize(nm);
        {
                const size_t otab = off(h.material_name, nm * 8);
                for (size_t m = 0; ok && (m < nm); m++) {
                        uint64_t pn;
                        std::memcpy(&pn, blob + otab + 8 * m, 8);
                        const size_t on = off(pn, 1);
                        if (!ok) break;
                        hp.material_name[m] = std::string(
                            reinterpret_cast<const char *>(blob + on),
                            strnlen(reinterpret_cast<const char *>(blob + on),
                                size - on));
                }
        }
        hp.material_I.resize(nm);
        if (ok)
                std::memcpy(hp.material_I.data(),
                    blob + off(h.material_I, nm * 8), nm * 8);
        hp.n_energy_loss_header = h.n_energy_loss_header;

        /* Build the image: blob + derived arrays. */
        const size_t o_blob = 0;
        const size_t o_coff = detail::align16((size_t)size);
        const size_t o_cel = detail::align16(o_coff + (nm + 1) * 4);
        const size_t o_cfr = detail::align16(o_cel + nc * 4);
        const size_t o_elt = detail::align16(o_cfr + nc * 8);
        /* the per cell records of the lean step (pb_hot_tables.h), 256 byte
         * aligned: a record never straddles a 128 byte line pair */
        const size_t n_cells = (ne > 0) ? (size_t)(ne - 1) : 0;
        const size_t o_hotK =
            (detail::align16(o_elt + nel * sizeof(Element)) + 255) &
            ~(size_t)255;
        const size_t o_hotX = o_hotK + nm * n_cells * HOTK_WORDS * sizeof(double);
        const size_t o_csdaK = detail::align16(
            o_hotX + nm * n_cells * HOTX_WORDS * sizeof(double));
        const size_t o_csdaX =
            o_csdaK + nm * n_cells * CSDAK_WORDS * sizeof(double);
        const size_t o_loc = detail::align16(
            o_csdaX + nm * n_cells * CSDAX_WORDS * sizeof(double));
        /* (the key ranges of the locators: from the tables in the blob) */
        int locK_min = 0, locK_keys = 1, locX_min = 0, locX_keys = 1;
        if (ok && (ne > 1)) {
                const double * K = reinterpret_cast<const double *>(
                    blob + off(h.table_K, ne * sizeof(double)));
                const double * X = reinterpret_cast<const double *>(blob +
                    off(h.table_X, N_SCHEMES * nm * ne * sizeof(double)));
                if (ok) {
                        locator_range(K, (int)ne, &locK_min, &locK_keys);
                        int lo = 0, hi = 0;
                        for (size_t m = 0; m < nm; m++) {
                                int kmin, nk;
                                locator_range(X + m * ne, (int)ne, &kmin, &nk);
                                if ((m == 0) || (kmin < lo)) lo = kmin;
                                if ((m == 0) || (kmin + nk > hi)) hi = kmin + nk;
                        }
                        locX_min = lo;
                        locX_keys = hi - lo;
                }
        }
        const size_t o_locK = o_loc;
        const size_t o_locX =
            detail::align16(o_locK + (size_t)locK_keys * sizeof(unsigned short));
        const size_t total = detail::align16(
            o_locX + nm * (size_t)locX_keys * sizeof(unsigned short));
        hp.image.assign(total, 0);
        std::memcpy(hp.image.data() + o_blob, blob, size);
        std::memcpy(hp.image.data() + o_coff, comp_offset.data(), (nm + 1) * 4);
        if (ok) {
                std::memcpy(hp.image.data() + o_cel, comp_element.data(), nc * 4);
                std::memcpy(
                    hp.image.data() + o_cfr, comp_fraction.data(), nc * 8);
        }
        std::memcpy(
            hp.image.data() + o_elt, elements.data(), nel * sizeof(Element));

        Physics & v = hp.view;
        std::memset(&v, 0, sizeof v);
        v.n_energies = h.n_energies;
        v.n_materials = h.n_materials;
        v.n_composites = h.n_composites;
        v.n_elements = h.n_elements;
        v.n_components = h.n_components;
        v.max_components = h.max_components;
        v.n_table_dcs = h.n_table_dcs;
        v.particle = h.particle;
        for (int i = 0; i < 3; i++) {
                v.model[i] = hp.model_index[i];
                v.user_dcs[i] = nullptr;
        }
        v.pad_ = 0;
        v.ctau = h.ctau;
        v.mass = h.mass;
        v.cutoff = h.cutoff;
        v.elastic_ratio = h.elastic_ratio;
        {
                volatile double one = 1., half = 0.5, m16 = -1. / 6.,
                                t13 = 1. / 3., m13 = -1. / 3.;
                v.c_exp_1 = exp(one);
                v.c_exp_1_2 = exp(half);
                v.c_exp_m1_6 = exp(m16);
                v.c_exp_1_3 = exp(t13);
                v.c_exp_m1_3 = exp(m13);
        }

        hp.relocs.clear();
        const unsigned char * pv = reinterpret_cast<const unsigned char *>(&v);
#define PB_RELOC(member, offset)                                               \
        hp.relocs.push_back(std::make_pair(                                    \
            (size_t)(reinterpret_cast<const unsigned char *>(&v.member) - pv), \
            (size_t)(offset)))
        const size_t d8 = sizeof(double), f4 = sizeof(float);
        PB_RELOC(K, off(h.table_K, ne * d8));
        PB_RELOC(K_dX, off(h.table_K_dX, N_SCHEMES * nm * ne * d8));
        PB_RELOC(K_dNI_in, off(h.table_K_dNI_in, nm * ne * d8));
        PB_RELOC(K_dNI_el, off(h.table_K_dNI_el, N_SCHEMES * nm * ne * d8));
        PB_RELOC(X, off(h.table_X, N_SCHEMES * nm * ne * d8));
        PB_RELOC(X_dK, off(h.table_X_dK, N_SCHEMES * nm * ne * d8));
        PB_RELOC(X_dT, off(h.table_X_dT, N_SCHEMES * nm * ne * d8));
        PB_RELOC(T, off(h.table_T, N_SCHEMES * nm * ne * d8));
        PB_RELOC(T_dK, off(h.table_T_dK, N_SCHEMES * nm * ne * d8));
        PB_RELOC(dE, off(h.table_dE, N_SCHEMES * nm * ne * d8));
        PB_RELOC(dE_dK, off(h.table_dE_dK, N_SCHEMES * nm * ne * d8));
        PB_RELOC(Omega, off(h.table_Omega, nm * ne * d8));
        PB_RELOC(Omega_dK, off(h.table_Omega_dK, nm * ne * d8));
        PB_RELOC(NI_el, off(h.table_NI_el, N_SCHEMES * nm * ne * d8));
        PB_RELOC(NI_el_dK, off(h.table_NI_el_dK, N_SCHEMES * nm * ne * d8));
        PB_RELOC(NI_in, off(h.table_NI_in, nm * ne * d8));
        PB_RELOC(NI_in_dK, off(h.table_NI_in_dK, nm * ne * d8));
        PB_RELOC(CS, off(h.table_CS, nm * ne * d8));
        PB_RELOC(CS_dK, off(h.table_CS_dK, nm * ne * d8));
        PB_RELOC(CSf, off(h.table_CSf, N_DEL_PROCESSES * nc * ne * d8));
        PB_RELOC(CSf_dK, off(h.table_CSf_dK, N_DEL_PROCESSES * nc * ne * d8));
        PB_RELOC(CSn, off(h.table_CSn, N_DEL_PROCESSES * nel * ne * d8));
        PB_RELOC(CSn_dK, off(h.table_CSn_dK, N_DEL_PROCESSES * nel * ne * d8));
        PB_RELOC(DCS,
            off(h.table_DCS,
                (N_DEL_PROCESSES - 1) * nel * ne * 2 * h.n_table_dcs * f4));
        PB_RELOC(DCS_x, off(h.table_DCS_x, h.n_table_dcs * f4));
        PB_RELOC(DCS_envelope,
            off(h.table_DCS_envelope,
                (N_DEL_PROCESSES - 1) * nel * ne * 2 * f4));
        PB_RELOC(Xt, off(h.table_Xt, N_DEL_PROCESSES * nel * ne * d8));
        PB_RELOC(Kt, off(h.table_Kt, nm * d8));
        PB_RELOC(Li, off(h.table_Li, nm * (N_LARMOR_ORDERS + 1) * ne * d8));
        PB_RELOC(
            Li_dK, off(h.table_Li_dK, nm * (N_LARMOR_ORDERS + 1) * ne * d8));
        PB_RELOC(a_max, off(h.table_a_max, nm * d8));
        PB_RELOC(b_max, off(h.table_b_max, N_SCHEMES * nm * d8));
        PB_RELOC(Mu0, off(h.table_Mu0, nm * ne * d8));
        PB_RELOC(Mu0_dK, off(h.table_Mu0_dK, nm * ne * d8));
        PB_RELOC(Lb, off(h.table_Lb, nm * ne * d8));
        PB_RELOC(Lb_dK, off(h.table_Lb_dK, nm * ne * d8));
        PB_RELOC(Ms1, off(h.table_Ms1, (N_SCHEMES + 1) * nm * ne * d8));
        PB_RELOC(Ms1_dK, off(h.table_Ms1_dK, (N_SCHEMES + 1) * nm * ne * d8));
        PB_RELOC(elements_in, off(h.elements_in, nm * 4));
        PB_RELOC(material_density, off(h.material_density, nm * d8));
        PB_RELOC(comp_offset, o_coff);
        PB_RELOC(comp_element, o_cel);
        PB_RELOC(comp_fraction, o_cfr);
        PB_RELOC(element, o_elt);
        PB_RELOC(hotK, o_hotK);
        PB_RELOC(hotX, o_hotX);
        PB_RELOC(csdaK, o_csdaK);
        PB_RELOC(csdaX, o_csdaX);
        PB_RELOC(locK, o_locK);
        PB_RELOC(locX, o_locX);
        v.locK_min = locK_min;
        v.locK_keys = locK_keys;
        v.locX_min = locX_min;
        v.locX_keys = locX_keys;
#undef PB_RELOC
        if (!ok) {
                err = "corrupted binary dump (pointer outside of the blob)";
                return 11;
        }
        hp.view = hp.view_at(hp.image.data());
        hot_tables_build(hp.view,
            reinterpret_cast<double *>(hp.image.data() + o_hotK),
            reinterpret_cast<double *>(hp.image.data() + o_hotX));
        csda_tables_build(hp.view,
            reinterpret_cast<double *>(hp.image.data() + o_csdaK),
            reinterpret_cast<double *>(hp.image.data() + o_csdaX));
        locator_build(hp.view.K, (int)ne, locK_min, locK_keys,
            reinterpret_cast<unsigned short *>(hp.image.data() + o_locK));
        for (size_t m = 0; m < nm; m++)
                locator_build(col_sm(hp.view, hp.view.X, MODE_CSDA, (int)m),
                    (int)ne, locX_min, locX_keys,
                    reinterpret_cast<unsigned short *>(
                        hp.image.data() + o_locX) +
                        m * (size_t)locX_keys);
        return 0;
}

inline int load_physics_stream(FILE * stream, HostPhysics & hp, std::string & err)
{
        int32_t head[2];
        if (fread(head, sizeof(int32_t), 2, stream) != 2) {
                err = "could not read the dump header";
                return 10;
        }
        if (head[0] != PHYSICS_BINARY_DUMP_TAG) {
                err = "incompatible version of binary dump";
                return 11;
        }
        if (head[1] < (int)sizeof(DumpHeader)) {
                err = "truncated binary dump";
                return 10;
        }
        std::vector<unsigned char> bytes(8 + (size_t)head[1]);
        std::memcpy(bytes.data(), head, 8);
        if (fread(bytes.data() + 8, head[1], 1, stream) != 1) {
                err = "truncated binary dump";
                return 10;
        }
        return load_physics_bytes(bytes.data(), bytes.size(), hp, err);
}

} /* namespace pb */

#endif
