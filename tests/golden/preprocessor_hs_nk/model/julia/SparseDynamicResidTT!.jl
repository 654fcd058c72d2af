function SparseDynamicResidTT!(T::Vector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real}, steady_state::Vector{<: Real})
@inbounds begin
    T[1] = -y[13];
    T[2] = 1/params[1];
    T[3] = -y[10];
    T[4] = -y[13] + y[9];
    T[5] = 1/(params[10]/400 + 1);
    T[6] = 1 - params[2];
    T[7] = params[11]/100;
    T[8] = params[12]/100;
    T[9] = params[13]/100;
    T[10] = T[6]*params[4];
    T[11] = -T[2];
end
    return nothing
end
