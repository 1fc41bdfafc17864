// Reference-side helper: turn the casadi::SX constraint function of a LoopConstraint::GenericImplicit into
// the grbda_phi_op program that grbda_schedule carries (include/grbda_cuda.h).
//
// The reference stores phi as a lambda over casadi::SX (`SymPhiFcn phi_sym_`,
// include/grbda/Dynamics/ClusterJoints/GenericJoint.h:15,49) and turns it into
// casadi::Function("phi", {q}, {phi(q)}) in its constructor (src/Dynamics/ClusterJoints/GenericJoint.cpp:45-55).
// An SX Function IS a straight-line program: CasADi exposes it instruction by instruction
// (Function::n_instructions / instruction_id / instruction_input / instruction_output /
// instruction_constant, casadi/core/function.hpp), over a work vector of scalar slots. walkSXFunction()
// copies that program into grbda_phi_op[] — no symbolic processing, no code generation.
//
// The walker is a template over the function type so that it can be compiled and tested without CasADi
// (tests/cpp/test_phi_bridge.cpp drives it with a stand-in that implements the same six members); with
// CasADi, instantiate it with casadi::Function and casadi's own opcode values:
//
//     #include <casadi/casadi.hpp>
//     #include <grbda_cuda_casadi.hpp>
//     casadi::SX q = casadi::SX::sym("q", n);                 // as GenericJoint.cpp:38-43
//     DVec<casadi::SX> q_vec(n);  casadi::copy(q, q_vec);
//     DVec<casadi::SX> phi_vec = lc.phi_sym_(JointCoordinate<casadi::SX>(q_vec, true));
//     casadi::SX phi(casadi::Sparsity::dense(phi_vec.rows(), 1));  casadi::copy(phi_vec, phi);
//     casadi::Function f("phi", {q}, {phi});
//     grbda_bridge::PhiProgram prog = grbda_bridge::walkSXFunction(f, grbda_bridge::casadiOpcodes());
//
// (INTEGRATION.md shows the complete ClusterTreeModel -> grbda_schedule walk around it.)
#ifndef GRBDA_CUDA_CASADI_HPP
#define GRBDA_CUDA_CASADI_HPP

#include <stdexcept>
#include <string>
#include <vector>
#include "grbda_cuda.h"

namespace grbda_bridge
{
    // opcode values of the function type being walked (casadi::Operation, casadi/core/calculus.hpp)
    struct Opcodes
    {
        int op_const, op_input, op_output, op_add, op_sub, op_mul, op_div, op_neg, op_sin, op_cos, op_sq, op_twice;
    };

#ifdef CASADI_CASADI_HPP
    inline Opcodes casadiOpcodes()
    {
        return Opcodes{casadi::OP_CONST, casadi::OP_INPUT, casadi::OP_OUTPUT, casadi::OP_ADD, casadi::OP_SUB,
                       casadi::OP_MUL,   casadi::OP_DIV,   casadi::OP_NEG,    casadi::OP_SIN, casadi::OP_COS,
                       casadi::OP_SQ,    casadi::OP_TWICE};
    }
#endif

    struct PhiProgram
    {
        std::vector<grbda_phi_op> ops;
        std::vector<int32_t> outputs; // op index of each constraint row
    };

    // grbda_phi_op::op values (include/grbda_cuda.h)
    enum
    {
        PHI_CONST = 0,
        PHI_INPUT = 1,
        PHI_ADD = 2,
        PHI_SUB = 3,
        PHI_MUL = 4,
        PHI_DIV = 5,
        PHI_NEG = 6,
        PHI_SIN = 7,
        PHI_COS = 8
    };

    // Fn needs: sz_w(), n_instructions(), instruction_id(k), instruction_input(k), instruction_output(k),
    // instruction_constant(k) with casadi::Function's meaning:
    //   OP_INPUT : instruction_input = {argument index, nonzero index}, instruction_output = {work slot}
    //   OP_OUTPUT: instruction_input = {work slot}, instruction_output = {result index, nonzero index}
    //   OP_CONST : instruction_constant = value, instruction_output = {work slot}
    //   others   : instruction_input = {work slot[, work slot]}, instruction_output = {work slot}
    // The function must have one argument (the spanning positions q) and one dense result (phi).
    template <typename Fn>
    PhiProgram walkSXFunction(const Fn &f, const Opcodes &oc)
    {
        PhiProgram p;
        std::vector<int32_t> slot((size_t)f.sz_w(), -1); // work slot -> op that produced its current value
        auto emit = [&](int32_t op, int32_t a, int32_t b, double val) {
            p.ops.push_back(grbda_phi_op{op, a, b, val});
            return (int32_t)p.ops.size() - 1;
        };
        auto value = [&](long long s) {
            if (s < 0 || (size_t)s >= slot.size() || slot[(size_t)s] < 0)
                throw std::runtime_error("walkSXFunction: instruction reads an unset work slot");
            return slot[(size_t)s];
        };
        for (long long k = 0; k < (long long)f.n_instructions(); k++)
        {
            const int id = (int)f.instruction_id(k);
            const auto in = f.instruction_input(k);
            const auto out = f.instruction_output(k);
            if (id == oc.op_output)
            {
                if (out.at(0) != 0)
                    throw std::runtime_error("walkSXFunction: phi must be the only result");
                const size_t row = (size_t)out.at(1);
                if (p.outputs.size() <= row)
                    p.outputs.resize(row + 1, -1);
                p.outputs[row] = value(in.at(0));
                continue;
            }
            int32_t r;
            if (id == oc.op_const)
                r = emit(PHI_CONST, -1, -1, (double)f.instruction_constant(k));
            else if (id == oc.op_input)
            {
                if (in.at(0) != 0)
                    throw std::runtime_error("walkSXFunction: phi must take the spanning positions as its only argument");
                r = emit(PHI_INPUT, -1, (int32_t)in.at(1), 0.0);
            }
            else if (id == oc.op_add)
                r = emit(PHI_ADD, value(in.at(0)), value(in.at(1)), 0.0);
            else if (id == oc.op_sub)
                r = emit(PHI_SUB, value(in.at(0)), value(in.at(1)), 0.0);
            else if (id == oc.op_mul)
                r = emit(PHI_MUL, value(in.at(0)), value(in.at(1)), 0.0);
            else if (id == oc.op_div)
                r = emit(PHI_DIV, value(in.at(0)), value(in.at(1)), 0.0);
            else if (id == oc.op_neg)
                r = emit(PHI_NEG, value(in.at(0)), -1, 0.0);
            else if (id == oc.op_sin)
                r = emit(PHI_SIN, value(in.at(0)), -1, 0.0);
            else if (id == oc.op_cos)
                r = emit(PHI_COS, value(in.at(0)), -1, 0.0);
            else if (id == oc.op_sq) // x^2
                r = emit(PHI_MUL, value(in.at(0)), value(in.at(0)), 0.0);
            else if (id == oc.op_twice) // 2 x
                r = emit(PHI_ADD, value(in.at(0)), value(in.at(0)), 0.0);
            else
                throw std::runtime_error("walkSXFunction: operation " + std::to_string(id) +
                                         " has no grbda_phi_op counterpart (supported: + - * / neg sin cos sq twice)");
            slot.at((size_t)out.at(0)) = r;
        }
        for (int32_t o : p.outputs)
            if (o < 0)
                throw std::runtime_error("walkSXFunction: a row of phi is structurally zero (give phi as a dense vector)");
        return p;
    }
} // namespace grbda_bridge

#endif // GRBDA_CUDA_CASADI_HPP
