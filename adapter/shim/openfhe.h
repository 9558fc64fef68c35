// MINIMAL lbcrypto SHIM — test infrastructure for adapter/BatchedFHEHIPPIE_b200.cpp, NOT OpenFHE.
//
// OpenFHE (openfheorg/openfhe-development) is not installed in this image, so the adapter that a maintainer
// drops next to the reference's sources is compiled here against this header instead: it declares ONLY the
// lbcrypto members the adapter touches, with the names and shapes OpenFHE 1.0.x has as recalled (every one to be
// checked against the installed headers: core/lattice/hal/default/dcrtpoly.h, core/math/hal/intnat/*,
// pke/cryptocontext.h, pke/scheme/bfvrns/cryptoparameters-bfvrns.h, pke/key/evalkeyrelin.h).  The classes are
// concrete and backed by std::vector<uint64_t>, so the adapter can also be RUN end to end (adapter/test_adapter.cpp
// fills them with ciphertexts encrypted by the oracle).  Nothing here is copied from OpenFHE.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef uint32_t usint;

namespace lbcrypto {

enum Format { EVALUATION = 0, COEFFICIENT = 1 };
enum MultiplicationTechnique { BEHZ = 0, HPS = 1, HPSPOVERQ = 2, HPSPOVERQLEVELED = 3 };
enum KeySwitchTechnique { INVALID_KS_TECH = 0, BV = 1, HYBRID = 2 };
struct SerType {
    enum SERBINARY { BINARY };
};

// NativeInteger: one machine word (sizeof == 8 is what lets the adapter hand limb vectors to the GPU without a copy)
class NativeInteger {
    uint64_t m_value;

   public:
    NativeInteger(uint64_t v = 0) : m_value(v) {}
    uint64_t ConvertToInt() const { return m_value; }
};
static_assert(sizeof(NativeInteger) == sizeof(uint64_t), "NativeInteger must be one word");

class NativeVector {
    std::vector<NativeInteger> m_data;
    NativeInteger m_modulus;

   public:
    NativeVector() = default;
    NativeVector(usint length, const NativeInteger& modulus) : m_data(length), m_modulus(modulus) {}
    usint GetLength() const { return (usint)m_data.size(); }
    const NativeInteger& GetModulus() const { return m_modulus; }
    NativeInteger& operator[](size_t i) { return m_data[i]; }
    const NativeInteger& operator[](size_t i) const { return m_data[i]; }
};

class ILNativeParams {
    usint m_order;
    NativeInteger m_modulus, m_root;

   public:
    ILNativeParams(usint order, NativeInteger modulus, NativeInteger root) : m_order(order), m_modulus(modulus), m_root(root) {}
    const NativeInteger& GetModulus() const { return m_modulus; }
    const NativeInteger& GetRootOfUnity() const { return m_root; }
    usint GetRingDimension() const { return m_order / 2; }
    usint GetCyclotomicOrder() const { return m_order; }
};

class NativePoly {
    std::shared_ptr<ILNativeParams> m_params;
    NativeVector m_values;
    Format m_format = EVALUATION;

   public:
    NativePoly() = default;
    NativePoly(std::shared_ptr<ILNativeParams> params, Format format, bool initializeElementToZero = false)
        : m_params(std::move(params)), m_values(m_params->GetRingDimension(), m_params->GetModulus()), m_format(format) {}
    const NativeVector& GetValues() const { return m_values; }
    void SetValues(NativeVector&& values, Format format) {
        m_values = std::move(values);
        m_format = format;
    }
    const NativeInteger& GetModulus() const { return m_params->GetModulus(); }
    usint GetLength() const { return m_values.GetLength(); }
    Format GetFormat() const { return m_format; }
    const std::shared_ptr<ILNativeParams>& GetParams() const { return m_params; }
};

struct BigInteger {};  // only a template tag here

template <typename IntType>
class ILDCRTParams {
    usint m_order;
    std::vector<std::shared_ptr<ILNativeParams>> m_params;

   public:
    ILDCRTParams(usint order, std::vector<std::shared_ptr<ILNativeParams>> towers) : m_order(order), m_params(std::move(towers)) {}
    const std::vector<std::shared_ptr<ILNativeParams>>& GetParams() const { return m_params; }
    usint GetRingDimension() const { return m_order / 2; }
    usint GetCyclotomicOrder() const { return m_order; }
};

class DCRTPoly {
   public:
    typedef ILDCRTParams<BigInteger> Params;

   private:
    std::shared_ptr<Params> m_params;
    std::vector<NativePoly> m_vectors;
    Format m_format = EVALUATION;

   public:
    DCRTPoly() = default;
    DCRTPoly(const std::shared_ptr<Params>& params, Format format, bool initializeElementToZero = false)
        : m_params(params), m_format(format) {
        for (const auto& t : params->GetParams()) m_vectors.emplace_back(t, format, initializeElementToZero);
    }
    usint GetNumOfElements() const { return (usint)m_vectors.size(); }
    const NativePoly& GetElementAtIndex(usint i) const { return m_vectors[i]; }
    void SetElementAtIndex(usint i, NativePoly&& element) { m_vectors[i] = std::move(element); }
    const std::shared_ptr<Params>& GetParams() const { return m_params; }
    Format GetFormat() const { return m_format; }
    usint GetRingDimension() const { return m_params->GetRingDimension(); }
};

class EncodingParamsImpl {
    uint64_t m_t, m_root;

   public:
    EncodingParamsImpl(uint64_t t, uint64_t root) : m_t(t), m_root(root) {}
    uint64_t GetPlaintextModulus() const { return m_t; }
    uint64_t GetPlaintextRootOfUnity() const { return m_root; }  // set by PackedEncoding::SetParams on first use
};
typedef std::shared_ptr<EncodingParamsImpl> EncodingParams;

template <class Element>
class CryptoParametersBase {
   public:
    virtual ~CryptoParametersBase() = default;
    virtual uint64_t GetPlaintextModulus() const = 0;
};

// the BFVrns members the adapter reads (names as recalled from cryptoparameters-bfvrns.h / -rns.h, 1.0.x)
class CryptoParametersBFVRNS : public CryptoParametersBase<DCRTPoly> {
   public:
    std::shared_ptr<DCRTPoly::Params> elementParams, paramsRl;
    EncodingParams encodingParams;
    MultiplicationTechnique multTech = HPSPOVERQ;
    KeySwitchTechnique ksTech = BV;
    usint digitSize = 0;
    // double tables of the HPS family (the adapter copies them when PSI_ADAPTER_TABLES_FROM_OPENFHE is defined)
    std::vector<double> qInv, rInv;
    std::vector<std::vector<double>> tQlSlHatInvModsDivsFrac;

    uint64_t GetPlaintextModulus() const override { return encodingParams->GetPlaintextModulus(); }
    const std::shared_ptr<DCRTPoly::Params>& GetElementParams() const { return elementParams; }
    const std::shared_ptr<DCRTPoly::Params>& GetParamsRl(usint l = 0) const { return paramsRl; }
    const EncodingParams& GetEncodingParams() const { return encodingParams; }
    MultiplicationTechnique GetMultiplicationTechnique() const { return multTech; }
    KeySwitchTechnique GetKeySwitchTechnique() const { return ksTech; }
    usint GetDigitSize() const { return digitSize; }
    const std::vector<double>& GetqInv() const { return qInv; }
    const std::vector<double>& GetrInv() const { return rInv; }
    const std::vector<double>& GettQlSlHatInvModsDivsFrac(usint l = 0) const { return tQlSlHatInvModsDivsFrac[l]; }
};

template <class Element>
class CryptoContextImpl;
template <class Element>
using CryptoContext = std::shared_ptr<CryptoContextImpl<Element>>;

class PlaintextImpl {};
typedef std::shared_ptr<PlaintextImpl> Plaintext;

template <class Element>
class CiphertextImpl {
    CryptoContext<Element> m_cc;
    std::string m_keyTag;
    std::vector<Element> m_elements;

   public:
    CiphertextImpl() = default;
    CiphertextImpl(CryptoContext<Element> cc, std::string keyTag) : m_cc(std::move(cc)), m_keyTag(std::move(keyTag)) {}
    const std::vector<Element>& GetElements() const { return m_elements; }
    void SetElements(std::vector<Element>&& elements) { m_elements = std::move(elements); }
    std::shared_ptr<CiphertextImpl<Element>> CloneEmpty() const { return std::make_shared<CiphertextImpl<Element>>(m_cc, m_keyTag); }
    const std::string& GetKeyTag() const { return m_keyTag; }
    CryptoContext<Element> GetCryptoContext() const { return m_cc; }
};
template <class Element>
using Ciphertext = std::shared_ptr<CiphertextImpl<Element>>;

template <class Element>
class PublicKeyImpl {
    std::string m_keyTag;

   public:
    explicit PublicKeyImpl(std::string tag = "") : m_keyTag(std::move(tag)) {}
    const std::string& GetKeyTag() const { return m_keyTag; }
};
template <class Element>
using PublicKey = std::shared_ptr<PublicKeyImpl<Element>>;

// relinearisation key, BV: A and B vectors with one DCRTPoly per digit
template <class Element>
class EvalKeyImpl {
   public:
    virtual ~EvalKeyImpl() = default;
    virtual const std::vector<Element>& GetAVector() const = 0;
    virtual const std::vector<Element>& GetBVector() const = 0;
};
template <class Element>
class EvalKeyRelinImpl : public EvalKeyImpl<Element> {
    std::vector<Element> m_a, m_b;

   public:
    void SetAVector(std::vector<Element>&& a) { m_a = std::move(a); }
    void SetBVector(std::vector<Element>&& b) { m_b = std::move(b); }
    const std::vector<Element>& GetAVector() const override { return m_a; }
    const std::vector<Element>& GetBVector() const override { return m_b; }
};
template <class Element>
using EvalKey = std::shared_ptr<EvalKeyImpl<Element>>;

template <class Element>
class CryptoContextImpl {
    std::shared_ptr<CryptoParametersBase<Element>> m_params;
    static std::map<std::string, std::vector<EvalKey<Element>>>& keyMap() {
        static std::map<std::string, std::vector<EvalKey<Element>>> m;
        return m;
    }

   public:
    explicit CryptoContextImpl(std::shared_ptr<CryptoParametersBase<Element>> p) : m_params(std::move(p)) {}
    const std::shared_ptr<CryptoParametersBase<Element>>& GetCryptoParameters() const { return m_params; }
    usint GetRingDimension() const {
        return std::dynamic_pointer_cast<CryptoParametersBFVRNS>(m_params)->GetElementParams()->GetRingDimension();
    }
    // static in OpenFHE too: the keys live in a process-wide map indexed by key tag
    static const std::vector<EvalKey<Element>>& GetEvalMultKeyVector(const std::string& keyTag) {
        auto it = keyMap().find(keyTag);
        if (it == keyMap().end()) throw std::runtime_error("no EvalMult key for this key tag");
        return it->second;
    }
    static void InsertEvalMultKey(const std::vector<EvalKey<Element>>& keys, const std::string& keyTag) { keyMap()[keyTag] = keys; }
    // EvalSum and rotation keys share one process-wide map in OpenFHE: automorphism index -> key, per key tag
    // (filled by DeserializeEvalSumKey / DeserializeEvalAutomorphismKey, SimpleFHEPSIServer.cpp:45-62)
    static std::map<usint, EvalKey<Element>>& GetEvalAutomorphismKeyMap(const std::string& keyTag) {
        auto it = autoKeyMap().find(keyTag);
        if (it == autoKeyMap().end()) throw std::runtime_error("no automorphism keys for this key tag");
        return *it->second;
    }
    static void InsertEvalAutomorphismKey(std::shared_ptr<std::map<usint, EvalKey<Element>>> keys, const std::string& keyTag) {
        autoKeyMap()[keyTag] = std::move(keys);
    }

   private:
    static std::map<std::string, std::shared_ptr<std::map<usint, EvalKey<Element>>>>& autoKeyMap() {
        static std::map<std::string, std::shared_ptr<std::map<usint, EvalKey<Element>>>> m;
        return m;
    }
};

}  // namespace lbcrypto
